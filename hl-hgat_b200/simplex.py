"""Per-mini-batch device-side operator tables.

The reference hands every layer an int64 COO `(edge_index, edge_weight)` (L0 / L1) and a sparse
COO boundary matrix `par` and re-derives what it needs on every call (gather/scatter over the
COO in `propagate`, `par.abs()` re-coalescing in `NodeEdgeInt`, lib/Hodge_Cheb_Conv.py:294-295,
494).  Here each operator is bucketed ONCE per mini-batch into int32 CSR tables (forward and
transpose) by `hl_csr_from_coo`, cached on the identity of the incoming tensors, and every
layer of the model reuses them.
"""
import weakref

import torch

from . import _native as N


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def csr_from_coo(row, col, val, nrows, tie=N.HL_TIE_POSITION, want_perm=False, row_is_float=False):
    """Bucket COO entries by `row` (int64, or float32 cluster ids when row_is_float) on the GPU.
    Returns (rowptr[int32, nrows+1], colidx[int32, nnz], vals[f32, nnz] | None, perm | None).
    Entries with row outside [0, nrows) (or +inf) are dropped; the outputs keep the full nnz
    allocation and rowptr[-1] says how many are valid."""
    L = N.lib()
    dev = row.device
    if not row.is_cuda:
        raise N.HlError("csr_from_coo needs CUDA tensors (no CPU fallback)")
    nnz = int(row.numel())
    row = row.contiguous()
    if row_is_float:
        assert row.dtype == torch.float32
    else:
        assert row.dtype == torch.int64
    col_c = None if col is None else col.contiguous()
    assert col_c is None or col_c.dtype == torch.int64
    val_c = None if val is None else val.contiguous()
    rowptr = torch.empty(nrows + 1, dtype=torch.int32, device=dev)
    colidx = torch.empty(nnz, dtype=torch.int32, device=dev)
    vals = torch.empty(nnz, dtype=torch.float32, device=dev) if val is not None else None
    perm = torch.empty(nnz, dtype=torch.int32, device=dev) if want_perm else None
    ws_bytes = L.hl_csr_from_coo_workspace(nnz, nrows)
    ws = _workspace(ws_bytes, dev)
    N.check(L.hl_csr_from_coo(None if row_is_float else row.data_ptr(), row.data_ptr() if row_is_float else None,
                              N.ptr(col_c), N.ptr(val_c), nnz, nrows, tie,
                              rowptr.data_ptr(), colidx.data_ptr(), N.ptr(vals), N.ptr(perm),
                              ws.data_ptr(), ws_bytes, N.stream_ptr()), "hl_csr_from_coo")
    return rowptr, colidx, vals, perm


class CsrOperator:
    """A = the propagate operator of `(edge_index, edge_weight)`: (A v)[i] = sum_{e: ei[1][e]=i}
    w[e] v[ei[0][e]] (lib/Hodge_Cheb_Conv.py:518-519, flow source->target).  Holds CSR by target
    row in COO order (forward) and CSR of A^T (backward), built lazily."""

    def __init__(self, edge_index, edge_weight, nrows):
        if edge_weight is None:
            edge_weight = torch.ones(edge_index.shape[1], dtype=torch.float32, device=edge_index.device)
        self.nrows = int(nrows)
        self.nnz = int(edge_index.shape[1])
        self._ei, self._ew = edge_index, edge_weight
        self._fwd = self._bwd = None
        self.factored = None            # optional Hodge1Factor: L1 = diag(s) B1^T B1 applied without the CSR

    @classmethod
    def from_csr(cls, rowptr, colidx, vals, nrows, symmetric=True, transpose=None):
        """Wrap CSR tables built on the device (hl_laplacian_fill): for a symmetric operator the same
        tables serve the forward and the adjoint pass."""
        op = cls.__new__(cls)
        op.nrows, op.nnz = int(nrows), int(colidx.numel())
        op._ei = op._ew = None
        op.factored = None
        op._fwd = (rowptr, colidx, vals)
        op._bwd = op._fwd if symmetric else transpose
        return op

    @property
    def fwd(self):
        if self._fwd is None:
            self._fwd = csr_from_coo(self._ei[1], self._ei[0], self._ew, self.nrows)[:3]
        return self._fwd

    @property
    def bwd(self):
        if self._bwd is None:
            self._bwd = csr_from_coo(self._ei[0], self._ei[1], self._ew, self.nrows)[:3]
        return self._bwd


class Hodge1Factor:
    """Factored form of an edge operator that IS the Hodge 1-Laplacian of the batch: L1 = diag(edge_scale) B1^T B1
    with edge_scale[e] = 2 / lambda_max(graph of e) (lib/Hodge_Dataset.py:456).  Built by the GPU constructor
    (which knows) or, for reference-format batches, from the operator's own diagonal (L1[e,e] = 2 edge_scale[e])
    when the caller opts in (functional.enable_factored_hodge1)."""

    def __init__(self, incidence, edge_scale):
        self.incidence, self.edge_scale = incidence, edge_scale.contiguous()

    @classmethod
    def from_operator(cls, op, incidence):
        rowptr, col, val = op.fwd
        counts = (rowptr[1:] - rowptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(op.nrows, device=col.device), counts, output_size=int(col.numel()))
        diag = torch.where(col.long() == rows, val, torch.zeros_like(val))
        scale = torch.zeros(op.nrows, dtype=torch.float32, device=col.device).index_add_(0, rows, diag) * 0.5
        return cls(incidence, scale)


class Incidence:
    """Tables of |B1| for a batch: `tail`/`head` (int32 [E], tail = edge_index[0]) and the
    node -> incident-edge CSR in ascending edge id (the order of the coalesced `par.abs()`)."""

    def __init__(self, edge_index, num_nodes):
        ei = edge_index
        e = int(ei.shape[1])
        self.num_nodes, self.num_edges = int(num_nodes), e
        self.tail = ei[0].to(torch.int32).contiguous()
        self.head = ei[1].to(torch.int32).contiguous()
        # entries interleaved (tail[e], e), (head[e], e) for e = 0, 1, ...: position order inside a node's row IS ascending
        # edge id, so the cheap position-stable bucketing gives the column order of the coalesced par.abs()
        ar = torch.arange(e, dtype=torch.int64, device=ei.device)
        self.rowptr, self.edge, _, _ = csr_from_coo(ei.t().reshape(-1), torch.repeat_interleave(ar, 2), None,
                                                    self.num_nodes, tie=N.HL_TIE_POSITION)

    @classmethod
    def from_tables(cls, tail, head, rowptr, edge, num_nodes):
        inc = cls.__new__(cls)
        inc.num_nodes, inc.num_edges = int(num_nodes), int(tail.numel())
        inc.tail, inc.head, inc.rowptr, inc.edge = tail, head, rowptr, edge
        return inc

    def degree(self):
        return (self.rowptr[1:] - self.rowptr[:-1]).to(torch.float32)


# ---------------------------------------------------------------------------------------------
# per-tensor caches: the unchanged reference models re-pass the same COO tensors to every layer and
# rebuild B1 at every stage (lib/Hodge_ST_Model.py:623-624); conversion happens once per mini-batch.
# The built tables hang off the SOURCE tensor (`edge_index._hl_ops`), so they live exactly as long as
# the batch they belong to -- nothing global keeps operators of earlier mini-batches resident in HBM.
# ---------------------------------------------------------------------------------------------
_ATTACHED = weakref.WeakSet()        # tensors carrying an `_hl_ops` table (for clear_caches)


def _table(t):
    tab = getattr(t, "_hl_ops", None)
    if tab is None:
        tab = t._hl_ops = {}
        _ATTACHED.add(t)
    return tab


def _sig(t):
    return None if t is None else (t.data_ptr(), tuple(t.shape), t._version)


def operator_for(edge_index, edge_weight, nrows):
    tab = _table(edge_index)
    key = ("op", nrows, _sig(edge_index), _sig(edge_weight))
    op = tab.get(key)
    if op is None:
        tab.clear()                   # an in-place update of the COO (new `_version`) invalidates what was built from it
        op = tab[key] = CsrOperator(edge_index, edge_weight, nrows)
    return op


def incidence_for(edge_index, num_nodes):
    tab = _table(edge_index)
    key = ("inc", num_nodes, _sig(edge_index))
    inc = tab.get(key)
    if inc is None:
        for k in [k for k in tab if k[0] == "inc"]:
            del tab[k]
        inc = tab[key] = Incidence(edge_index, num_nodes)
    return inc


def clear_caches():
    """Forget every table built so far (the static buffers of a captured step change content between replays, and the
    rebuild has to be part of the capture)."""
    for t in list(_ATTACHED):
        t._hl_ops.clear()
