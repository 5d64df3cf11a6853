"""GPU simplex-graph construction for a whole mini-batch.

Replaces the reference's per-graph CPU pipeline (lib/Hodge_Dataset.py:447-456,467-468: dense B1,
dense B1 B1^T, dense eigh, dense_to_sparse) and the block-diagonal collation (:40-48) with sort /
segment kernels over the concatenated directed edge list of all graphs; nothing dense is formed."""
import torch

from . import _native as N
from .simplex import CsrOperator, Hodge1Factor, Incidence, csr_from_coo


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=dev)


class SimplexBatch:
    """Everything the model needs about one mini-batch of graphs, resident on the device."""

    def coo(self, side):
        """The reference's format: int64 COO `edge_index_{t,s}` [2,nnz] (row-major) + `edge_weight_{t,s}`."""
        rowptr, col, val = (self.op_t if side == "t" else self.op_s).fwd
        counts = (rowptr[1:] - rowptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(counts.numel(), device=col.device), counts)
        return torch.stack([rows, col.long()]), val


def build_simplex_batch(src, dst, node_counts, edge_attr=None, max_steps=None):
    """src/dst: int64 [M] directed edges with GLOBAL node ids (graph g owns the contiguous id range
    given by node_counts); node_counts: int64 [G].  Returns a SimplexBatch."""
    L = N.lib()
    dev = src.device
    if not src.is_cuda:
        raise N.HlError("build_simplex_batch needs CUDA tensors (no CPU fallback)")
    src, dst = src.contiguous(), dst.contiguous()
    node_counts = node_counts.to(dev, torch.int64)
    G = int(node_counts.numel())
    node_ptr64 = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    node_ptr64[1:] = torch.cumsum(node_counts, 0)
    n_nodes = int(node_ptr64[-1])
    m = int(src.numel())
    st = N.stream_ptr()

    # 1. unique undirected edges, lexicographic
    tail = torch.empty(m, dtype=torch.int32, device=dev)
    head = torch.empty(m, dtype=torch.int32, device=dev)
    attr_out = torch.empty(m, dtype=torch.int64, device=dev) if edge_attr is not None else None
    n_edges_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    nb = L.hl_build_edges_workspace(m)
    ws = _ws(nb, dev)
    N.check(L.hl_build_edges(src.data_ptr(), dst.data_ptr(), m, n_nodes,
                             N.ptr(None if edge_attr is None else edge_attr.contiguous()),
                             tail.data_ptr(), head.data_ptr(), N.ptr(attr_out), n_edges_dev.data_ptr(),
                             ws.data_ptr(), nb, st), "hl_build_edges")
    n_edges = int(n_edges_dev)                                  # the one host sync of the construction
    tail, head = tail[:n_edges].contiguous(), head[:n_edges].contiguous()

    return _finish(tail, head, node_counts, node_ptr64, n_nodes, attr_out, max_steps, lexicographic=True)


def simplex_batch_from_edges(tail, head, node_counts, max_steps=None):
    """Operators of a batch whose UNIQUE undirected edges are already known: `tail` < `head` (int32, global node
    ids), in ANY order -- the order is kept (rows of L1 follow it), as the coarse graphs of MLGC need, whose edge
    list is in first-appearance order (lib/Hodge_Dataset.py:262-275)."""
    dev = tail.device
    node_counts = node_counts.to(dev, torch.int64)
    node_ptr64 = torch.zeros(node_counts.numel() + 1, dtype=torch.int64, device=dev)
    node_ptr64[1:] = torch.cumsum(node_counts, 0)
    return _finish(tail.to(torch.int32).contiguous(), head.to(torch.int32).contiguous(), node_counts, node_ptr64,
                   int(node_ptr64[-1]), None, max_steps, lexicographic=False)


def _finish(tail, head, node_counts, node_ptr64, n_nodes, attr_out, max_steps, lexicographic):
    L = N.lib()
    dev = tail.device
    st = N.stream_ptr()
    G = int(node_counts.numel())
    n_edges = int(tail.numel())
    # 2. node -> incident edges (ascending edge id)
    ar = torch.arange(n_edges, dtype=torch.int64, device=dev)
    rows = torch.cat([tail, head]).long()
    inc_rowptr, inc_edge, _, _ = csr_from_coo(rows, torch.cat([ar, ar]), None, n_nodes, tie=N.HL_TIE_COLUMN)
    inc = Incidence.from_tables(tail, head, inc_rowptr, inc_edge, n_nodes)

    # 3. lambda_max per graph
    node_ptr = node_ptr64.to(torch.int32)
    max_nodes = max(int(node_counts.max()) if G else 1, 1)
    steps = max_nodes if max_steps is None else max(1, min(max_steps, max_nodes))
    lam = torch.empty(G, dtype=torch.float32, device=dev)
    change = torch.empty(G, dtype=torch.float32, device=dev)
    nb = L.hl_lambda_max_workspace(G, max_nodes, steps)
    ws = _ws(nb, dev)
    N.check(L.hl_lambda_max(node_ptr.data_ptr(), G, max_nodes, inc_rowptr.data_ptr(), inc_edge.data_ptr(),
                            tail.data_ptr(), head.data_ptr(), steps, lam.data_ptr(), change.data_ptr(),
                            ws.data_ptr(), nb, st), "hl_lambda_max")

    # 4. Laplacians
    node_graph = torch.repeat_interleave(torch.arange(G, device=dev, dtype=torch.int32), node_counts,
                                         output_size=n_nodes)
    r0 = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev)
    r1 = torch.empty(n_edges + 1, dtype=torch.int32, device=dev)
    nb = L.hl_laplacian_rowptr_workspace(n_edges, n_nodes)
    ws = _ws(nb, dev)
    N.check(L.hl_laplacian_rowptr(tail.data_ptr(), head.data_ptr(), n_edges, n_nodes, inc_rowptr.data_ptr(),
                                  r0.data_ptr(), r1.data_ptr(), ws.data_ptr(), nb, st), "hl_laplacian_rowptr")
    nnz0, nnz1 = int(r0[-1]), int(r1[-1])
    c0 = torch.empty(nnz0, dtype=torch.int32, device=dev)
    v0 = torch.empty(nnz0, dtype=torch.float32, device=dev)
    c1 = torch.empty(nnz1, dtype=torch.int32, device=dev)
    v1 = torch.empty(nnz1, dtype=torch.float32, device=dev)
    N.check(L.hl_laplacian_fill(tail.data_ptr(), head.data_ptr(), n_edges, n_nodes, inc_rowptr.data_ptr(),
                                inc_edge.data_ptr(), node_graph.data_ptr(), lam.data_ptr(),
                                r0.data_ptr(), c0.data_ptr(), v0.data_ptr(), r1.data_ptr(), c1.data_ptr(), v1.data_ptr(),
                                st), "hl_laplacian_fill")

    if not lexicographic and nnz0:
        # hl_laplacian_fill emits L0 neighbours in incident-edge order, which is ascending only for a
        # lexicographic edge list; restore the row-major ascending order of dense_to_sparse
        rows0 = torch.repeat_interleave(torch.arange(n_nodes, device=dev), (r0[1:] - r0[:-1]).long(), output_size=nnz0)
        perm = torch.argsort(rows0 * n_nodes + c0.long())
        c0, v0 = c0[perm].contiguous(), v0[perm].contiguous()

    b = SimplexBatch()
    b.num_graphs, b.num_nodes, b.num_edges = G, n_nodes, n_edges
    b.tail, b.head = tail, head
    b.edge_index = torch.stack([tail.long(), head.long()])
    b.edge_attr = None if attr_out is None else attr_out[:n_edges]
    b.incidence = inc
    b.lambda_max, b.lambda_last_change = lam, change
    b.op_t = CsrOperator.from_csr(r0, c0, v0, n_nodes)
    b.op_s = CsrOperator.from_csr(r1, c1, v1, n_edges)
    if n_edges:                         # the constructor knows that op_s = diag(2/lambda) B1^T B1: keep the factored form too
        b.op_s.factored = Hodge1Factor(inc, 2.0 / lam[node_graph[tail.long()].long()])
    b.num_node1 = node_counts
    b.num_edge1 = torch.bincount(node_graph[tail.long()].long(), minlength=G) if n_edges else torch.zeros_like(node_counts)
    b.node_graph = node_graph
    return b
