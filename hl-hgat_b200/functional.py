"""Differentiable host-side wrappers (torch.autograd.Function) over the C ABI.

Forward and backward both run hand-written sm_100a kernels from libhlhgat.so, including the dense Theta / MLP
transforms and their data / weight / bias gradients (tcgen05 3xTF32); cuBLAS (through torch.mm) is only the
fallback for dense shapes the tensor-core kernels do not take.
"""
import contextlib
import ctypes as C
import math

import torch

from . import _native as N
from . import lanes as _lanes
from .simplex import CsrOperator, Hodge1Factor, Incidence, csr_from_coo  # noqa: F401

_FAMILY = {"laguerre": N.HL_LAGUERRE, "cheb": N.HL_CHEB}


def _side(csr, nrows, x=None, ld_x=0, t=None, ld_t=0, t_stride=0, g0=None, ld_g0=0):
    s = N.ConvSide()
    s.rowptr, s.colidx, s.vals = csr[0].data_ptr(), csr[1].data_ptr(), csr[2].data_ptr()
    s.nrows = nrows
    s.nnz_hint = min(int(csr[1].numel()), 2 ** 31 - 1)
    s.x, s.ld_x = N.ptr(x), ld_x
    s.t, s.ld_t, s.t_stride = N.ptr(t), ld_t, t_stride
    s.g0, s.ld_g0 = N.ptr(g0), ld_g0
    return s


_FACTORED = {"enabled": False}


def enable_factored_hodge1(flag=True):
    """Opt in to the factored application L1 x = diag(s) B1^T (B1 x) for operators that carry a `Hodge1Factor`
    (operators built by construct.build_simplex_batch, or attached by the model classes from the batch's boundary
    matrix): 4 row gathers per edge instead of one per nonzero of L1.  Off by default -- the CSR path reproduces the
    reference's summation order bit for bit, the factored one only to fp32 rounding (~1e-7)."""
    _FACTORED["enabled"] = bool(flag)


def factored_hodge1_enabled():
    return _FACTORED["enabled"]


def _use_factored(op):
    return _FACTORED["enabled"] and getattr(op, "factored", None) is not None


def _hodge1_struct(op):
    f = op.factored
    inc = f.incidence
    h = N.Hodge1Operator()
    h.inc_rowptr, h.inc_edge, h.tail, h.head = inc.rowptr.data_ptr(), inc.edge.data_ptr(), inc.tail.data_ptr(), inc.head.data_ptr()
    h.edge_scale, h.n_nodes, h.n_edges = f.edge_scale.data_ptr(), inc.num_nodes, inc.num_edges
    return h


def poly_basis_fwd(family, K, ops, xs, width):
    """T_1..T_{K-1} for several operators in K-1 shared launches.  Returns stacked [K-1,R,width]."""
    L = N.lib()
    csr = [i for i, op in enumerate(ops) if not _use_factored(op)]
    sides = (N.ConvSide * max(len(csr), 1))()
    outs = [None] * len(ops)
    for i, (op, x) in enumerate(zip(ops, xs)):
        x2, ldx = N.row_major(x)
        t = torch.empty((max(K - 1, 0), op.nrows, width), dtype=torch.float32, device=x.device)
        outs[i] = t
        if i in csr:
            sides[csr.index(i)] = _side(op.fwd, op.nrows, x=x2, ld_x=ldx, t=t, ld_t=width, t_stride=op.nrows * width)
        elif K > 1:
            tmp = torch.empty((op.factored.incidence.num_nodes, width), dtype=torch.float32, device=x.device)
            N.check(L.hl_poly_basis_hodge1_fwd(family, K, _hodge1_struct(op), x2.data_ptr(), ldx, t.data_ptr(), width,
                                               op.nrows * width, tmp.data_ptr(), width, N.stream_ptr()), "hl_poly_basis_hodge1_fwd")
    if K > 1 and csr:
        N.check(L.hl_poly_basis_fwd(family, K, sides, len(csr), width, N.stream_ptr()), "hl_poly_basis_fwd")
    return outs


def poly_basis_bwd(family, K, ops, g0s, gts, width):
    L = N.lib()
    csr = [i for i, op in enumerate(ops) if not _use_factored(op)]
    sides = (N.ConvSide * max(len(csr), 1))()
    for i, (op, g0, gt) in enumerate(zip(ops, g0s, gts)):
        if i in csr:
            sides[csr.index(i)] = _side(op.bwd, op.nrows, t=gt, ld_t=width, t_stride=op.nrows * width, g0=g0, ld_g0=width)
        elif K > 1:
            tmp = torch.empty((op.factored.incidence.num_nodes, width), dtype=torch.float32, device=g0.device)
            N.check(L.hl_poly_basis_hodge1_bwd(family, K, _hodge1_struct(op), g0.data_ptr(), width, gt.data_ptr(), width,
                                               op.nrows * width, tmp.data_ptr(), width, N.stream_ptr()), "hl_poly_basis_hodge1_bwd")
    if K > 1 and csr:
        N.check(L.hl_poly_basis_bwd(family, K, sides, len(csr), width, N.stream_ptr()), "hl_poly_basis_bwd")


import os as _os
_EXPERIMENT_SKIP_WGRAD = _os.environ.get("HL_EXPERIMENT_SKIP_WGRAD") == "1"
_WGRAD2 = _os.environ.get("HL_WGRAD2", "1") != "0"          # two weight gradients sharing g in one launch (A/B switch)
_GEMM_MODE = {"tensor": True}


def set_dense_backend(name):
    """"tcgen05" (default): 3xTF32 tensor-core GEMM from libhlhgat; "cublas": torch.mm (fp32 SIMT SGEMM)."""
    _GEMM_MODE["tensor"] = name == "tcgen05"

_BACKWARD_MODE = {"accumulate": False}


class accumulate_into_grads:
    """Fused gradient accumulation.  Wrap `loss.backward()` in this context (training.GraphedTrainStep does):
    inside it the kernels that produce weight / bias / BatchNorm-affine gradients ADD them straight into an
    existing dense fp32 `param.grad` (e.g. the views of parallel.FlatGradBucket) and hand autograd `None`
    for those inputs -- the same values autograd's own `grad.add_(new)` would leave there, without one extra
    elementwise launch per parameter.  Outside the context (and under `torch.autograd.grad`) gradients are
    returned as usual.

    With lanes enabled these kernels run on auxiliary streams and write into `param.grad` behind autograd's back
    (its end-of-backward stream synchronisation does not cover them), so leaving the context JOINS the lanes:
    after `with accumulate_into_grads(): loss.backward()` the current stream is ordered after every gradient
    write and an all-reduce / optimizer step may follow directly."""

    def __enter__(self):
        self.prev = _BACKWARD_MODE["accumulate"]
        _BACKWARD_MODE["accumulate"] = True
        return self

    def __exit__(self, *a):
        _BACKWARD_MODE["accumulate"] = self.prev
        if not self.prev:
            _lanes.join()                      # free when nothing was issued on a side stream


def _grad_target(p):
    """`p.grad` if the producing kernel may accumulate into it directly, else None."""
    if not _BACKWARD_MODE["accumulate"] or not isinstance(p, torch.nn.Parameter):
        return None
    g = p.grad
    if g is None or not g.is_cuda or g.dtype != torch.float32 or g.shape != p.shape or g.stride(-1) != 1:
        return None
    if g.dim() == 2 and g.stride(0) < g.shape[1]:
        return None
    return g


def _wgrad_lane(fused, *tensors):
    """Context for issuing a weight / bias gradient: the lane's auxiliary stream when the result is accumulated
    in place into `param.grad` (nothing in the autograd graph waits for it; lanes.weight_grad_lane), else the
    current stream."""
    return _lanes.weight_grad_lane(*tensors) if fused else contextlib.nullcontext()


class WeightSplitPlan:
    """Takes the tf32 hi / lo split of the weights off the critical chain of a training step.

    Every tensor-core GEMM needs its weight operand pre-split (`hl_tf32_split`, a few microseconds, ~130 launches per
    ZINC step), and weights change only in the optimizer step.  Inside `with plan:` the first pass RECORDS every
    split request (which parameter view, transposed or not, packed with which second view) and keeps its hi / lo
    buffers; every later `with plan:` re-issues all recorded splits up front on the plan's own stream -- one more
    parallel branch of the whole-step graph, overlapping the CSR bucketing -- and the GEMM call sites just wait for
    its event and read the persistent buffers.  Parameter storage must stay in place (in-place optimizer updates,
    `load_state_dict`); call `clear()` after anything that re-allocates parameters."""

    def __init__(self):
        self.entries = {}          # key -> (views kept alive, hi, lo, [hl_tf32_split argument tuples])
        self.stream = None
        self.ready = None
        self._table = None         # device copy of the descriptor table (rebuilt when entries were added)
        self._table_len = 0
        self._max_elements = 0

    def clear(self):
        self.entries.clear()
        self._table, self._table_len = None, 0

    def _build_table(self, device):
        descs = [a for _, _, _, launches in self.entries.values() for a in launches]
        arr = (N.SplitDesc * len(descs))()
        for d, (src, ld_src, rows, cols, tr, hi, lo, ld_out) in zip(arr, descs):
            d.src, d.hi, d.lo, d.ld_src, d.ld_out, d.rows, d.cols, d.transpose = src, hi, lo, ld_src, ld_out, rows, cols, tr
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._table = raw.to(device)
        self._table_len = len(descs)
        self._n_entries = len(self.entries)
        self._max_elements = max(r * c for _, _, r, c, *_ in descs)

    def __enter__(self):
        global _PLAN
        assert _PLAN is None, "WeightSplitPlan scopes do not nest"
        _PLAN = self
        self.ready = None
        if self.entries and self._table is not None and self._n_entries == len(self.entries):
            cur = torch.cuda.current_stream()
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=cur.device)
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                N.check(N.lib().hl_tf32_split_batch(self._table.data_ptr(), self._table_len, self._max_elements, N.stream_ptr()),
                        "hl_tf32_split_batch")
            self.ready = torch.cuda.Event()
            self.ready.record(self.stream)       # every GEMM call site waits for it (`get`), which also rejoins
        return self                              # the plan's stream into a graph capture

    def __exit__(self, *a):
        global _PLAN
        _PLAN = None
        if self.entries and (self._table is None or self._n_entries != len(self.entries)):
            if not torch.cuda.is_current_stream_capturing():       # the H2D copy of the table cannot be captured
                self._build_table(next(iter(self.entries.values()))[1].device)

    def get(self, key):
        ent = self.entries.get(key)
        if ent is None or self.ready is None:
            return None
        torch.cuda.current_stream().wait_event(self.ready)
        return ent[1], ent[2]


_PLAN = None


def _split_weights(key, views, shape, launches, zero):
    """hi / lo buffers of `shape` for a split request: from the active plan if it has them, else split now on the
    current stream (and hand the buffers to the plan, if one is recording).  `launches`: callables (hi, lo) ->
    hl_tf32_split argument tuple (without the stream)."""
    plan = _PLAN
    if plan is not None:
        hit = plan.get(key)
        if hit is not None:
            return hit
    dev = views[0].device
    alloc = torch.zeros if zero else torch.empty
    hi, lo = alloc(shape, dtype=torch.float32, device=dev), alloc(shape, dtype=torch.float32, device=dev)
    args = [mk(hi, lo) for mk in launches]
    L, st = N.lib(), N.stream_ptr()
    for a in args:
        N.check(L.hl_tf32_split(*a, st), "hl_tf32_split")
    if plan is not None and key not in plan.entries and not torch.cuda.is_current_stream_capturing():
        plan.entries[key] = (views, hi, lo, args)      # hi / lo allocated inside a capture belong to the graph's pool
    return hi, lo


def _view_key(t):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()))


class BnTiles:
    """Request for the BatchNorm statistics of a Linear / polynomial-conv output from the GEMM epilogue
    (hl_gemm2_bn_tf32x3): `with bn_stats_from_epilogue(nvalid) as req: h = linear(...)`, then `bn_act_train(h, ..., tiles=req)`.
    The launch that writes the FINAL values of the output fills `part` ([ceil(M/32)][2][N]: mean | M2 of every 32-row
    block over the rows below *nvalid) and `key`; when it could not (cuBLAS path, an odd shape) `part` stays None and
    the BatchNorm runs its own statistics pass as before."""
    __slots__ = ("nvalid", "part", "key", "out")

    def __init__(self, nvalid):
        self.nvalid, self.part, self.key, self.out = nvalid, None, None, None

    def matches(self, x, nvalid):
        return self.part is not None and self.key == (x.data_ptr(), x.shape[0], x.shape[1], x.stride(0)) \
            and N.ptr(self.nvalid) == N.ptr(nvalid)


_BN_EPILOGUE = _os.environ.get("HL_BN_EPILOGUE", "1") != "0"
_BN_REQ = [None]


class bn_stats_from_epilogue:
    def __init__(self, nvalid=None, enabled=True):
        self.req = BnTiles(nvalid) if (enabled and _BN_EPILOGUE) else None

    def __enter__(self):
        self.prev, _BN_REQ[0] = _BN_REQ[0], self.req
        return self.req

    def __exit__(self, *a):
        _BN_REQ[0] = self.prev


def _take_bn_request():
    """The pending request (at most one producer serves it: the first autograd Function whose forward runs inside the
    `with`)."""
    req, _BN_REQ[0] = _BN_REQ[0], None
    return req


def _bn_part_for(bn, M, n_out, device):
    if bn is None or M < 1:
        return None
    return torch.empty(N.lib().hl_gemm_bn_part_floats(M, n_out), dtype=torch.float32, device=device)


def _bn_done(bn, part, out):
    # `out` is held so that its address cannot be handed to another tensor of the same shape while the request lives
    bn.part, bn.key, bn.out = part, (out.data_ptr(), out.shape[0], out.shape[1], out.stride(0)), out


def dense(a, w, bias=None, out=None, accumulate=False, transpose_w=False, bn=None):
    """out (=|+=) a @ w.T (+ bias)   [transpose_w: a @ w].  fp32-accurate tcgen05 GEMM (3xTF32 split) when the
    shape allows (N % 16 == 0, 16-byte aligned rows), else the cuBLAS fp32 GEMM.  `bn` (BnTiles): also the
    BatchNorm block statistics of the result from the epilogue."""
    L = N.lib()
    M, K = a.shape
    n_out = w.shape[1] if transpose_w else w.shape[0]
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.float32, device=a.device)
        accumulate = False
    if _GEMM_MODE["tensor"] and M > 0 and n_out % 16 == 0 and a.stride(1) == 1 and a.stride(0) % 4 == 0 \
            and a.data_ptr() % 16 == 0 and K % 4 == 0 and out.stride(1) == 1:
        kp = K
        # w is [n_out, K] (or [K, n_out] when transpose_w): split into tf32-exact hi and remainder lo, [n_out, K]
        rows, cols = (K, n_out) if transpose_w else (n_out, K)
        tr = 1 if transpose_w else 0
        hi, lo = _split_weights(("1", tr) + _view_key(w), (w,), (n_out, kp),
                                [lambda h, l: (w.data_ptr(), w.stride(0), rows, cols, tr, h.data_ptr(), l.data_ptr(), kp)], False)
        part = _bn_part_for(bn, M, n_out, a.device)
        if part is None:
            rc = L.hl_gemm_tf32x3(a.data_ptr(), a.stride(0), hi.data_ptr(), lo.data_ptr(), kp, M, n_out, K, N.ptr(bias),
                                  out.data_ptr(), out.stride(0), 1 if accumulate else 0, N.stream_ptr())
        else:
            rc = L.hl_gemm2_bn_tf32x3(a.data_ptr(), a.stride(0), K, None, 0, 0, hi.data_ptr(), lo.data_ptr(), kp, M, n_out,
                                      N.ptr(bias), out.data_ptr(), out.stride(0), 1 if accumulate else 0, part.data_ptr(),
                                      N.ptr(bn.nvalid), N.stream_ptr())
        if rc == 0:
            if part is not None:
                _bn_done(bn, part, out)
            return out
        if rc != 1:
            N.check(rc, "hl_gemm_tf32x3")
    wt = w if transpose_w else w.t()
    if accumulate:
        out.addmm_(a, wt)
        if bias is not None:
            out.add_(bias)
    elif bias is not None:
        torch.addmm(bias, a, wt, out=out)
    else:
        torch.mm(a, wt, out=out)
    return out


def dense2(a1, w1, a2, w2, bias=None, out=None, accumulate=False, bn=None):
    """out (=|+=) a1 @ w1.T + a2 @ w2.T (+ bias) in ONE tensor-core launch (the sum stays in TMEM)."""
    L = N.lib()
    M, k1 = a1.shape
    k2 = a2.shape[1]
    n_out = w1.shape[0]
    ok = (_GEMM_MODE["tensor"] and M > 0 and n_out % 16 == 0 and k1 % 4 == 0 and k2 % 4 == 0
          and all(t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in (a1, a2)))
    if ok:
        k1p = (k1 + 31) // 32 * 32
        kt = k1p + k2
        st = N.stream_ptr()
        # packed [w1 | pad | w2] along K; the pad columns k1..k1p of the packed weights must be 0
        hi, lo = _split_weights(("2",) + _view_key(w1) + _view_key(w2), (w1, w2), (n_out, kt),
                                [lambda h, l: (w1.data_ptr(), w1.stride(0), n_out, k1, 0, h.data_ptr(), l.data_ptr(), kt),
                                 lambda h, l: (w2.data_ptr(), w2.stride(0), n_out, k2, 0, h[:, k1p:].data_ptr(), l[:, k1p:].data_ptr(), kt)],
                                k1p != k1)
        if out is None:
            out = torch.empty((M, n_out), dtype=torch.float32, device=a1.device)
            accumulate = False
        part = _bn_part_for(bn, M, n_out, a1.device)
        rc = L.hl_gemm2_bn_tf32x3(a1.data_ptr(), a1.stride(0), k1, a2.data_ptr(), a2.stride(0), k2, hi.data_ptr(), lo.data_ptr(), kt,
                                  M, n_out, N.ptr(bias), out.data_ptr(), out.stride(0), 1 if accumulate else 0, N.ptr(part),
                                  N.ptr(bn.nvalid) if part is not None else None, st)
        if rc == 0:
            if part is not None:
                _bn_done(bn, part, out)
            return out
        if rc != 1:
            N.check(rc, "hl_gemm2_bn_tf32x3")
    out = dense(a1, w1, bias, out=out, accumulate=accumulate)
    return dense(a2, w2, None, out=out, accumulate=True, bn=bn)


class WgradReducePlan:
    """Collects the split reduces of the weight gradients of a training step and runs them in ONE launch.

    A tensor-core weight gradient is two launches: the split-row pass (partial planes in a workspace) and a small
    reduce that sums the planes into the gradient (8-10 us each, ~36 per ZINC step, ~170 per TSP step).  Nothing reads
    a weight gradient before the all-reduce / optimizer, so inside `with plan:` a weight gradient that ACCUMULATES into
    its destination (`accumulate_into_grads`: the flat gradient bucket) only runs the first launch
    (`hl_wgrad_deferred_tf32x3`) and leaves a descriptor here; `flush()` -- after `lanes.join()`, before the all-reduce --
    sums all of them with `hl_wgrad_reduce_batch` (descriptors as kernel parameters: nothing to upload, capturable).
    Two gradients for the same destination are not batched together: the second one is reduced on the spot."""

    def __init__(self):
        self.descs, self.keep, self.targets = [], [], set()

    def __enter__(self):
        global _RPLAN
        assert _RPLAN is None, "WgradReducePlan scopes do not nest"
        _RPLAN = self
        return self

    def __exit__(self, *a):
        global _RPLAN
        _RPLAN = None
        if self.descs:
            self.flush()

    @staticmethod
    def _extent(t):
        rows, cols, ld = (1, t.shape[0], t.shape[0]) if t.dim() == 1 else (t.shape[0], t.shape[1], t.stride(0))
        return t.data_ptr(), rows, cols, ld

    @staticmethod
    def _overlap(a, b):
        if a[0] > b[0]:
            a, b = b, a
        (pa, ra, ca, la), (pb, rb, cb, lb) = a, b
        if pa + ((ra - 1) * la + ca) * 4 <= pb:
            return False                                          # disjoint address ranges
        if la == lb and (pb - pa) % 4 == 0:                       # column blocks of one matrix (gw[:, :d] and gw[:, d:])
            off = ((pb - pa) // 4) % la
            if off >= ca and off + cb <= la:
                return False
        return True

    def claim(self, *tensors):
        """True when no destination overlaps one that is already pending in this batch (then they are registered)."""
        ext = [self._extent(t) for t in tensors if t is not None]
        if any(self._overlap(e, f) for e in ext for f in self.targets):
            return False
        self.targets.update(ext)
        return True

    def add(self, desc, *keep):
        self.descs.append(desc)
        self.keep.append(keep)                     # workspace (and operands) stay allocated until the reduce has run

    def flush(self):
        """Sum every pending weight gradient on the CURRENT stream (the caller has joined the streams the split-row
        passes ran on)."""
        if self.descs:
            arr = (N.WgradReduceDesc * len(self.descs))(*self.descs)
            N.check(N.lib().hl_wgrad_reduce_batch(arr, len(self.descs), N.stream_ptr()), "hl_wgrad_reduce_batch")
            for keep in self.keep:                 # the allocator may hand the workspaces out again after this point
                for t in keep:
                    t.record_stream(torch.cuda.current_stream())
        self.descs, self.keep, self.targets = [], [], set()


_RPLAN = None
# Opt-in (HL_WGRAD_DEFER=1): measured SLOWER on the ZINC step (4.41 -> 4.50 ms, same box, two alternating runs each) and
# neutral on TSP (30.67 vs 30.53 ms): the per-gradient reduces run on the auxiliary streams under the backward pass and cost
# nothing there, while the batched launch sits in the serial tail between the last gradient and the all-reduce.
_WGRAD_DEFER = _os.environ.get("HL_WGRAD_DEFER", "0") == "1"


def _wgrad_deferred(g, ldg, x1, ldx1, x2, ldx2, out1, out2, bias_out, bias_accumulate):
    """The split-row pass only, reduce left to the active WgradReducePlan.  Returns None when the gradient cannot be
    deferred (no plan, shape outside the tensor-core path, destination already pending), else the rc of the launch
    (0: all done later, 2: bias gradient not folded)."""
    plan = _RPLAN
    if plan is None or not _WGRAD_DEFER or not _GEMM_MODE["tensor"]:
        return None
    L = N.lib()
    R, fo = g.shape
    fi = x1.shape[1]
    nb = L.hl_wgrad2_tf32x3_workspace(R, fo, fi) if x2 is not None else L.hl_wgrad_tf32x3_workspace(R, fo, fi)
    if nb == 0:
        return None
    if not plan.claim(out1, out2, bias_out):
        return None
    ws = torch.empty(nb, dtype=torch.uint8, device=g.device)
    desc = N.WgradReduceDesc()
    rc = L.hl_wgrad_deferred_tf32x3(g.data_ptr(), ldg, x1.data_ptr(), ldx1, N.ptr(x2), ldx2, R, fo, fi, out1.data_ptr(), out1.stride(0),
                                    N.ptr(out2), out2.stride(0) if out2 is not None else 0, 1, N.ptr(bias_out),
                                    1 if bias_accumulate else 0, ws.data_ptr(), nb, C.byref(desc), N.stream_ptr())
    if rc == 1:
        return None
    if rc not in (0, 2):
        N.check(rc, "hl_wgrad_deferred_tf32x3")
    plan.add(desc, ws)
    return rc


def wgrad(g, x, out=None, accumulate=False, bias_out=None, bias_accumulate=False):
    """dW[Fo,Fi] (=|+=) g[R,Fo]^T x[R,Fi] (deterministic split-row reduction); `out` may be a column slice.
    `bias_out` [Fo] (optional): also (=|+=) the column sums of g -- folded into the tensor-core launch when it can
    be (the converter warps that move g^T into tensor memory add up their column), else through hl_colsum."""
    L = N.lib()
    g, ldg = N.row_major(g)
    x, ldx = N.row_major(x)
    R, fo = g.shape
    fi = x.shape[1]
    if out is None:
        out = torch.empty((fo, fi), dtype=torch.float32, device=g.device)
        accumulate = False
    if _EXPERIMENT_SKIP_WGRAD:                   # timing experiment only (HL_EXPERIMENT_SKIP_WGRAD=1): gradients are garbage
        return out
    acc = 1 if accumulate else 0
    done = False
    if accumulate and (bias_out is None or bias_accumulate):
        rc = _wgrad_deferred(g, ldg, x, ldx, None, 0, out, None, bias_out, bias_accumulate)
        if rc is not None:
            if rc == 2 and bias_out is not None:
                _colsum_into(g, ldg, bias_out, bias_accumulate)
            return out
    if _GEMM_MODE["tensor"]:
        nb = L.hl_wgrad_tf32x3_workspace(R, fo, fi)
        ws = torch.empty(nb, dtype=torch.uint8, device=g.device)
        rc = L.hl_wgrad_bias_tf32x3(g.data_ptr(), ldg, x.data_ptr(), ldx, R, fo, fi, out.data_ptr(), out.stride(0), acc,
                                    N.ptr(bias_out), 1 if bias_accumulate else 0, ws.data_ptr(), nb, N.stream_ptr())
        if rc == 0 or rc == 2:
            done = True
            if rc == 0:
                bias_out = None                      # folded in
        elif rc != 1:
            N.check(rc, "hl_wgrad_bias_tf32x3")
    if not done:
        nb = L.hl_wgrad_workspace(R, fo, fi)
        ws = torch.empty(nb, dtype=torch.uint8, device=g.device)
        N.check(L.hl_wgrad(g.data_ptr(), ldg, x.data_ptr(), ldx, R, fo, fi, out.data_ptr(), out.stride(0), acc,
                           ws.data_ptr(), nb, N.stream_ptr()), "hl_wgrad")
    if bias_out is not None:
        _colsum_into(g, ldg, bias_out, bias_accumulate)
    return out


def wgrad2(g, x1, x2, out1, out2, accumulate=False, bias_out=None, bias_accumulate=False):
    """out1 (=|+=) g^T x1 and out2 (=|+=) g^T x2 for two activations of the same shape that share g: ONE tensor-core
    launch and one reduce (`hl_wgrad2_bias_tf32x3`) when the shape allows, else two `wgrad` calls."""
    L = N.lib()
    done = False
    if _GEMM_MODE["tensor"] and not _EXPERIMENT_SKIP_WGRAD and _WGRAD2 and x1.shape == x2.shape:
        g2, ldg = N.row_major(g)
        a, lda = N.row_major(x1)
        b, ldb = N.row_major(x2)
        R, fo = g2.shape
        fi = a.shape[1]
        if accumulate and (bias_out is None or bias_accumulate):
            rc = _wgrad_deferred(g2, ldg, a, lda, b, ldb, out1, out2, bias_out, bias_accumulate)
            if rc is not None:
                if rc == 2 and bias_out is not None:
                    _colsum_into(g2, ldg, bias_out, bias_accumulate)
                return out1, out2
        nb = L.hl_wgrad2_tf32x3_workspace(R, fo, fi)
        ws = torch.empty(nb, dtype=torch.uint8, device=g.device)
        acc = 1 if accumulate else 0
        rc = L.hl_wgrad2_bias_tf32x3(g2.data_ptr(), ldg, a.data_ptr(), lda, b.data_ptr(), ldb, R, fo, fi, out1.data_ptr(), out1.stride(0),
                                     out2.data_ptr(), out2.stride(0), acc, N.ptr(bias_out), 1 if bias_accumulate else 0,
                                     ws.data_ptr(), nb, N.stream_ptr())
        if rc == 0 or rc == 2:
            done = True
            if rc == 2 and bias_out is not None:
                _colsum_into(g2, ldg, bias_out, bias_accumulate)
        elif rc != 1:
            N.check(rc, "hl_wgrad2_bias_tf32x3")
    if not done:
        wgrad(g, x1, out1, accumulate=accumulate, bias_out=bias_out, bias_accumulate=bias_accumulate)
        wgrad(g, x2, out2, accumulate=accumulate)
    return out1, out2


def _colsum_into(g, ldg, out, accumulate):
    L = N.lib()
    R, f = g.shape
    nb = L.hl_colsum_workspace(R, f)
    ws = torch.empty(nb, dtype=torch.uint8, device=g.device)
    N.check(L.hl_colsum(g.data_ptr(), ldg, R, f, out.data_ptr(), 1 if accumulate else 0, ws.data_ptr(), nb, N.stream_ptr()), "hl_colsum")


def colsum(g, out=None):
    """out[f] (=|+=) sum_r g[r,f]; accumulates when `out` is given."""
    L = N.lib()
    g, ldg = N.row_major(g)
    R, f = g.shape
    acc = 0 if out is None else 1
    if out is None:
        out = torch.empty(f, dtype=torch.float32, device=g.device)
    nb = L.hl_colsum_workspace(R, f)
    ws = torch.empty(nb, dtype=torch.uint8, device=g.device)
    N.check(L.hl_colsum(g.data_ptr(), ldg, R, f, out.data_ptr(), acc, ws.data_ptr(), nb, N.stream_ptr()), "hl_colsum")
    return out


class _Linear(torch.autograd.Function):
    """y = [xa | xb] W^T + b without materialising the concat (xb optional); dgrad on cuBLAS, weight and
    bias gradients through hl_wgrad / hl_colsum."""

    @staticmethod
    def forward(ctx, xa, xb, weight, bias):
        N.require_cuda_f32(xa, xb, weight, bias)
        d = xa.shape[1]
        xa = xa.contiguous() if xa.stride(1) != 1 else xa
        bn = _take_bn_request()
        if xb is None:
            y = dense(xa, weight, bias, bn=bn)
        else:
            xb = xb.contiguous() if xb.stride(1) != 1 else xb
            y = dense2(xa, weight[:, :d], xb, weight[:, d:], bias, bn=bn)
        ctx.save_for_backward(xa, xb, weight)
        ctx.has_bias = bias is not None
        ctx.params = (weight, bias)                 # the Parameter objects (for fused gradient accumulation)
        return y

    @staticmethod
    def backward(ctx, g):
        xa, xb, weight = ctx.saved_tensors
        g = g.contiguous()
        d = xa.shape[1]
        ga = gb = gw = gbias = None
        want_bias = ctx.has_bias and ctx.needs_input_grad[3]
        tgt_b = _grad_target(ctx.params[1]) if want_bias else None
        if ctx.needs_input_grad[2]:
            tgt = _grad_target(ctx.params[0])
            fold = want_bias and (tgt is None) == (tgt_b is None)      # bias gradient rides on the first weight-gradient launch
            if fold:
                gbias = tgt_b if tgt_b is not None else torch.empty(weight.shape[0], dtype=torch.float32, device=g.device)
            with _wgrad_lane(tgt is not None, g, xa, xb):
                gw = torch.empty_like(weight) if tgt is None else tgt
                if xb is not None and xb.shape == xa.shape:
                    wgrad2(g, xa, xb, gw[:, :d], gw[:, d:], accumulate=tgt is not None, bias_out=gbias if fold else None,
                           bias_accumulate=tgt_b is not None)
                else:
                    wgrad(g, xa, gw[:, :d], accumulate=tgt is not None, bias_out=gbias if fold else None,
                          bias_accumulate=tgt_b is not None)
                    if xb is not None:
                        wgrad(g, xb, gw[:, d:], accumulate=tgt is not None)
            if tgt is not None:
                gw = None
            if fold:
                want_bias = False
                if tgt_b is not None:
                    gbias = None
        if want_bias:
            with _wgrad_lane(tgt_b is not None, g):
                gbias = colsum(g, out=tgt_b)
            if tgt_b is not None:
                gbias = None
        if ctx.needs_input_grad[0]:
            ga = dense(g, weight[:, :d], transpose_w=True)
        if xb is not None and ctx.needs_input_grad[1]:
            gb = dense(g, weight[:, d:], transpose_w=True)
        return ga, gb, gw, gbias


def linear(x, weight, bias=None, x2=None):
    return _Linear.apply(x, x2, weight, bias)


class _LinearPart(torch.autograd.Function):
    """y = x W[:, c0:c1]^T (+ b) (+ addend): one column block of a Linear applied on its own, optionally
    accumulated IN PLACE onto `addend` (a fresh tensor nobody else reads, e.g. a transfer result).  Lets the
    NodeEdgeInt MLP project before it transfers: (1/D)|B1| (x_s Wa^T) + x_t Wb^T + b instead of
    [(1/D)|B1| x_s | x_t] W^T + b (lib/Hodge_Cheb_Conv.py:294, :307-308) -- the transfer then moves the layer width f
    instead of the dense-connection width d."""

    @staticmethod
    def forward(ctx, x, weight, bias, addend, c0, c1):
        N.require_cuda_f32(x, weight, bias, addend)
        x = x.contiguous() if x.stride(1) != 1 else x
        w = weight[:, c0:c1]
        if addend is None:
            y = dense(x, w, bias)
        else:
            y = dense(x, w, bias, out=addend, accumulate=True)
            ctx.mark_dirty(addend)
        ctx.save_for_backward(x, weight)
        ctx.cols = (c0, c1)
        ctx.params = (weight, bias)
        ctx.has_bias, ctx.has_addend = bias is not None, addend is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        c0, c1 = ctx.cols
        g = g.contiguous()
        gx = gw = gbias = None
        want_bias = ctx.has_bias and ctx.needs_input_grad[2]
        tgt_b = _grad_target(ctx.params[1]) if want_bias else None
        if ctx.needs_input_grad[1]:
            tgt = _grad_target(ctx.params[0])
            fold = want_bias and (tgt is None) == (tgt_b is None)
            if fold:
                gbias = tgt_b if tgt_b is not None else torch.empty(weight.shape[0], dtype=torch.float32, device=g.device)
            with _wgrad_lane(tgt is not None, g, x):
                gw = torch.zeros_like(weight) if tgt is None else tgt      # only the column block is written
                wgrad(g, x, gw[:, c0:c1], accumulate=tgt is not None, bias_out=gbias if fold else None,
                      bias_accumulate=tgt_b is not None)
            if tgt is not None:
                gw = None
            if fold:
                want_bias = False
                if tgt_b is not None:
                    gbias = None
        if want_bias:
            with _wgrad_lane(tgt_b is not None, g):
                gbias = colsum(g, out=tgt_b)
            if tgt_b is not None:
                gbias = None
        if ctx.needs_input_grad[0]:
            gx = dense(g, weight[:, c0:c1], transpose_w=True)
        return gx, gw, gbias, (g if ctx.has_addend else None), None, None


def linear_part(x, weight, c0, c1, bias=None, addend=None):
    return _LinearPart.apply(x, weight, bias, addend, int(c0), int(c1))


_PTT = {"enabled": False}


def enable_project_then_transfer(flag=True):
    """Opt in: NodeEdgeInt applies the column block of its first Linear that acts on the TRANSFERRED features before
    the transfer whenever the layer is narrower than the dense-connection buffer (results equal to fp32 rounding,
    not bit for bit: the summation order of the reference is [transfer, then Linear])."""
    _PTT["enabled"] = bool(flag)


def project_then_transfer_enabled():
    return _PTT["enabled"]


class _PolyConv(torch.autograd.Function):
    """out = sum_k T_k(x) W_k^T + b for one operator (lib/Hodge_Cheb_Conv.py:480-515 / :394-439).
    x is [R, width] (already flattened), inner = last-dim size the Linear layers act on."""

    @staticmethod
    def forward(ctx, x, bias, op, family, inner, *weights):
        N.require_cuda_f32(x, bias, *weights)
        K = len(weights)
        x = x.contiguous()
        R, width = x.shape
        (t,) = poly_basis_fwd(family, K, [op], [x], width)
        xv = x.view(-1, inner)
        bn = _take_bn_request()
        if inner != width:                          # [R, T, C] input: the BatchNorm rows are not the GEMM rows
            bn = None
        if K == 1:
            out = dense(xv, weights[0], bias, bn=bn)
        else:
            out = dense2(xv, weights[0], t[0].view(-1, inner), weights[1], bias, bn=bn if K == 2 else None)
            for k in range(2, K, 2):
                last = bn if k + 2 >= K else None   # statistics of the FINAL values: the launch that completes the sum
                if k + 1 < K:
                    dense2(t[k - 1].view(-1, inner), weights[k], t[k].view(-1, inner), weights[k + 1], None, out=out, accumulate=True,
                           bn=last)
                else:
                    dense(t[k - 1].view(-1, inner), weights[k], None, out=out, accumulate=True, bn=last)
        ctx.op, ctx.family, ctx.inner, ctx.has_bias = op, family, inner, bias is not None
        ctx.params = (bias, weights)
        ctx.save_for_backward(x, t, *weights)
        return out

    @staticmethod
    def backward(ctx, g):
        x, t, *weights = ctx.saved_tensors
        K, inner, op = len(weights), ctx.inner, ctx.op
        R, width = x.shape
        g = g.contiguous()
        need_x = ctx.needs_input_grad[0]
        want_bias = ctx.has_bias and ctx.needs_input_grad[1]
        tgt_b = _grad_target(ctx.params[0]) if want_bias else None
        gb = None
        gws = [None] * K
        k = 0
        while k < K:
            if not ctx.needs_input_grad[5 + k]:
                k += 1
                continue
            src = x if k == 0 else t[k - 1]
            tgt = _grad_target(ctx.params[1][k])
            fold = want_bias and (tgt is None) == (tgt_b is None) and g.shape[0] == src.numel() // inner
            if fold:                                   # bias gradient rides on this weight-gradient launch
                gb = tgt_b if tgt_b is not None else torch.empty(g.shape[1], dtype=torch.float32, device=g.device)
            # two consecutive orders share g: one launch for both weight gradients (hl_wgrad2_bias_tf32x3)
            pair = k + 1 < K and ctx.needs_input_grad[6 + k] and (_grad_target(ctx.params[1][k + 1]) is None) == (tgt is None)
            with _wgrad_lane(tgt is not None, g, src, t[k] if pair else None):
                if pair:
                    tgt2 = _grad_target(ctx.params[1][k + 1])
                    o1 = tgt if tgt is not None else torch.empty_like(weights[k])
                    o2 = tgt2 if tgt2 is not None else torch.empty_like(weights[k + 1])
                    wgrad2(g, src.view(-1, inner), t[k].view(-1, inner), o1, o2, accumulate=tgt is not None,
                           bias_out=gb if fold else None, bias_accumulate=tgt_b is not None)
                    gws[k], gws[k + 1] = (None, None) if tgt is not None else (o1, o2)
                else:
                    gw = wgrad(g, src.view(-1, inner), out=tgt, accumulate=tgt is not None, bias_out=gb if fold else None,
                               bias_accumulate=tgt_b is not None)
                    gws[k] = None if tgt is not None else gw
            if fold:
                want_bias = False
                if tgt_b is not None:
                    gb = None
            k += 2 if pair else 1
        if want_bias:
            with _wgrad_lane(tgt_b is not None, g):
                gb = colsum(g, out=tgt_b)
            if tgt_b is not None:
                gb = None
        gx = None
        if need_x:
            g0 = dense(g, weights[0], transpose_w=True).view(R, width)
            gt = torch.empty((max(K - 1, 0), R, width), dtype=torch.float32, device=x.device)
            for k in range(1, K):
                dense(g, weights[k], out=gt[k - 1].view(-1, inner), transpose_w=True)
            poly_basis_bwd(ctx.family, K, [op], [g0], [gt], width)
            gx = g0
        return (gx, gb, None, None, None, *gws)


def poly_conv(x, weights, bias, op, family="laguerre"):
    """x: [R,C] or [R,T,C]; weights: list of K [Fout,C] tensors; returns [R,Fout] / [R,T,Fout]."""
    shp = x.shape
    inner = shp[-1]
    out = _PolyConv.apply(x.reshape(shp[0], -1), bias, op, _FAMILY[family], inner, *weights)
    return out.view(*shp[:-1], -1)


# ---------------------------------------------------------------------------------------------
# single SpMM launches (used by the DEMO HodgeLaguerreFastConv quirk path)
# ---------------------------------------------------------------------------------------------
def poly_spmm(csr, nrows, xg, epi, c=(0.0, 0.0, 0.0, 0.0), p1=None, p2=None, p3=None, out=None):
    L = N.lib()
    xg, ldx = N.row_major(xg)
    width = xg.shape[1]
    if out is None:
        out = torch.empty((nrows, width), dtype=torch.float32, device=xg.device)
    P = (N.SpmmProblem * 1)()
    P[0].rowptr, P[0].colidx, P[0].vals, P[0].nrows = csr[0].data_ptr(), csr[1].data_ptr(), csr[2].data_ptr(), nrows
    P[0].nnz_hint = min(int(csr[1].numel()), 2 ** 31 - 1)
    P[0].xg, P[0].ld_xg = xg.data_ptr(), ldx
    for name, t in (("p1", p1), ("p2", p2), ("p3", p3)):
        if t is not None:
            t2, ld = N.row_major(t)
            setattr(P[0], name, t2.data_ptr())
            setattr(P[0], "ld_" + name, ld)
    P[0].out, P[0].ld_out = out.data_ptr(), out.stride(0)
    cc = (C.c_float * 4)(*c)
    N.check(L.hl_poly_spmm(P, 1, width, epi, cc, N.stream_ptr()), "hl_poly_spmm")
    return out


# ---------------------------------------------------------------------------------------------
# node <-> edge simplex transfer
# ---------------------------------------------------------------------------------------------
def _segment_reduce(rowptr, colidx, nrows, src, post, row_scale=None, cscale=1.0, src_scale=None, out=None):
    L = N.lib()
    src, ld = N.row_major(src)
    width = src.shape[1]
    if out is None:
        out = torch.empty((nrows, width), dtype=torch.float32, device=src.device)
    N.check(L.hl_segment_reduce(rowptr.data_ptr(), N.ptr(colidx), nrows, src.data_ptr(), ld, N.ptr(src_scale),
                                out.data_ptr(), out.stride(0), width, post, N.ptr(row_scale), cscale,
                                N.stream_ptr()), "hl_segment_reduce")
    return out


def _endpoint_gather(inc, src, node_rcp, cscale, out=None):
    L = N.lib()
    src, ld = N.row_major(src)
    width = src.shape[1]
    if out is None:
        out = torch.empty((inc.num_edges, width), dtype=torch.float32, device=src.device)
    N.check(L.hl_endpoint_gather(inc.tail.data_ptr(), inc.head.data_ptr(), inc.num_edges, src.data_ptr(), ld,
                                 N.ptr(node_rcp), out.data_ptr(), out.stride(0), width, cscale, N.stream_ptr()),
            "hl_endpoint_gather")
    return out


class _EdgeToNode(torch.autograd.Function):
    """x_s2t = (1/D) * (|B1| x_s)   (lib/Hodge_Cheb_Conv.py:294)."""

    @staticmethod
    def forward(ctx, x_s, D, inc):
        N.require_cuda_f32(x_s, D)
        ctx.inc = inc
        ctx.save_for_backward(D)
        return _segment_reduce(inc.rowptr, inc.edge, inc.num_nodes, x_s, N.HL_POST_RCP_ROW, row_scale=D)

    @staticmethod
    def backward(ctx, g):
        (D,) = ctx.saved_tensors
        return _endpoint_gather(ctx.inc, g, D, 1.0), None, None


class _NodeToEdge(torch.autograd.Function):
    """x_t2s = (|B1|^T x_t) / 2   (lib/Hodge_Cheb_Conv.py:295)."""

    @staticmethod
    def forward(ctx, x_t, inc):
        N.require_cuda_f32(x_t)
        ctx.inc = inc
        return _endpoint_gather(inc, x_t, None, 0.5)

    @staticmethod
    def backward(ctx, g):
        inc = ctx.inc
        return _segment_reduce(inc.rowptr, inc.edge, inc.num_nodes, g, N.HL_POST_CONST, cscale=0.5), None


def edge_to_node(x_s, D, inc):
    return _EdgeToNode.apply(x_s, D.contiguous(), inc)


def node_to_edge(x_t, inc):
    return _NodeToEdge.apply(x_t, inc)


class _BoundaryAbsDiff(torch.autograd.Function):
    """|B1^T x_t| / 2 per edge (lib/Hodge_ST_Model.py:848, the TSP readout)."""

    @staticmethod
    def forward(ctx, x_t, inc):
        N.require_cuda_f32(x_t)
        x_t, ld = N.row_major(x_t)
        width = x_t.shape[1]
        out = torch.empty((inc.num_edges, width), dtype=torch.float32, device=x_t.device)
        N.check(N.lib().hl_boundary_absdiff_fwd(inc.tail.data_ptr(), inc.head.data_ptr(), inc.num_edges, x_t.data_ptr(), ld,
                                                out.data_ptr(), out.stride(0), width, 0.5, N.stream_ptr()),
                "hl_boundary_absdiff_fwd")
        ctx.inc, ctx.ld = inc, ld
        ctx.save_for_backward(x_t)
        return out

    @staticmethod
    def backward(ctx, g):
        (x_t,) = ctx.saved_tensors
        inc = ctx.inc
        g, ldg = N.row_major(g)
        width = x_t.shape[1]
        dx = torch.empty((inc.num_nodes, width), dtype=torch.float32, device=g.device)
        N.check(N.lib().hl_boundary_absdiff_bwd(inc.rowptr.data_ptr(), inc.edge.data_ptr(), inc.tail.data_ptr(),
                                                inc.head.data_ptr(), inc.num_nodes, x_t.data_ptr(), ctx.ld,
                                                g.data_ptr(), ldg, dx.data_ptr(), dx.stride(0), width, 0.5, N.stream_ptr()),
                "hl_boundary_absdiff_bwd")
        return dx, None


def boundary_absdiff(x_t, inc):
    return _BoundaryAbsDiff.apply(x_t, inc)


# ---------------------------------------------------------------------------------------------
# attention gate, cluster pooling, per-graph readout
# ---------------------------------------------------------------------------------------------
class _AttGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qc, qs, k, lam, sigma):
        N.require_cuda_f32(qc, qs, k)
        qc, qs, k = qc.contiguous(), qs.contiguous(), k.contiguous()
        R, dk = k.shape
        a = torch.empty((R, 1), dtype=torch.float32, device=k.device)
        N.check(N.lib().hl_att_gate_fwd(qc.data_ptr(), qs.data_ptr(), k.data_ptr(), R, dk, lam, sigma,
                                        a.data_ptr(), N.stream_ptr()), "hl_att_gate_fwd")
        ctx.lam, ctx.sigma = lam, sigma
        ctx.save_for_backward(qc, qs, k, a)
        return a

    @staticmethod
    def backward(ctx, da):
        qc, qs, k, a = ctx.saved_tensors
        R, dk = k.shape
        dqc, dqs, dkk = torch.empty_like(qc), torch.empty_like(qs), torch.empty_like(k)
        N.check(N.lib().hl_att_gate_bwd(qc.data_ptr(), qs.data_ptr(), k.data_ptr(), a.data_ptr(),
                                        da.contiguous().data_ptr(), R, dk, ctx.lam, ctx.sigma,
                                        dqc.data_ptr(), dqs.data_ptr(), dkk.data_ptr(), N.stream_ptr()),
                "hl_att_gate_bwd")
        return dqc, dqs, dkk, None, None


def att_gate(q_cross, q_self, k, lam, sigma="sigmoid"):
    """sigma(((1-lam) <q_cross,k> + lam <q_self,k>)/sqrt(dk))   (lib/Hodge_Cheb_Conv.py:299-304)."""
    code = {"sigmoid": N.HL_SIGMA_SIGMOID, "relu": N.HL_SIGMA_RELU}[sigma]
    return _AttGate.apply(q_cross, q_self, k, float(lam), code)


class Segments:
    """Bucketing of source rows into output rows for a mean reduction: cluster ids (int64 or the
    reference's float ids with +inf = dropped) or contiguous per-graph counts."""

    def __init__(self, rowptr, members, owner, nrows, nsrc):
        self.rowptr, self.members, self.owner, self.nrows, self.nsrc = rowptr, members, owner, nrows, nsrc
        cnt = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(torch.float32)
        self.inv_count = 1.0 / cnt

    @classmethod
    def from_index(cls, index, nrows=None):
        idx = index.reshape(-1)
        is_float = idx.dtype.is_floating_point
        if nrows is None:
            finite = idx[torch.isfinite(idx)] if is_float else idx
            nrows = int(finite.max().item()) + 1 if finite.numel() else 0
        # ids the bucketing drops (+inf = edge inside a cluster, anything outside [0, nrows)) own nothing: -1, which
        # the adjoint (hl_owner_gather) turns into a zero gradient row instead of reading g[owner] out of bounds
        if is_float:
            idx = idx.to(torch.float32)
            keep = torch.isfinite(idx) & (idx >= 0) & (idx < nrows)
            owner = torch.where(keep, idx, torch.full_like(idx, -1.0)).to(torch.int32)
        else:
            idx = idx.to(torch.int64)
            keep = (idx >= 0) & (idx < nrows)
            owner = torch.where(keep, idx, torch.full_like(idx, -1)).to(torch.int32)
        rowptr, members, _, _ = csr_from_coo(idx, None, None, nrows, tie=N.HL_TIE_POSITION, row_is_float=is_float)
        return cls(rowptr, members, owner, nrows, idx.numel())

    @classmethod
    def from_counts(cls, counts, total=None, ghost_last=False):
        """Contiguous segments (sorted `batch` vector): counts[g] rows per graph.  Pass `total`
        (= counts.sum(), known on the host) to stay free of device->host syncs (CUDA-graph capture).
        ghost_last: the last segment is the ghost graph of a padded fixed-capacity batch (training.pad_batch) -- hundreds
        of all-zero rows that one lane group would walk serially; it is left EMPTY (its mean is the same zero row) and
        its rows own nothing."""
        counts = counts.to(torch.int64)
        ids = torch.arange(counts.numel(), device=counts.device, dtype=torch.int32)
        owner = torch.repeat_interleave(ids, counts, output_size=total)
        if ghost_last:                              # device-side only (captured into the step graph)
            counts = torch.where(ids == counts.numel() - 1, torch.zeros_like(counts), counts)
            owner = torch.where(owner == counts.numel() - 1, torch.full_like(owner, -1), owner)
        ptr = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=counts.device)
        ptr[1:] = torch.cumsum(counts, 0)
        return cls(ptr, None, owner, counts.numel(), int(owner.numel()))


class _SegmentMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, scale, seg):
        N.require_cuda_f32(src, scale)
        src = src.contiguous()
        sc = None if scale is None else scale.reshape(-1).contiguous()
        ctx.seg = seg
        ctx.save_for_backward(src, sc)
        return _segment_reduce(seg.rowptr, seg.members, seg.nrows, src, N.HL_POST_MEAN, src_scale=sc)

    @staticmethod
    def backward(ctx, g):
        src, sc = ctx.saved_tensors
        seg = ctx.seg
        g = g.contiguous()
        dsrc = torch.empty_like(src)
        dsc = torch.empty(src.shape[0], dtype=torch.float32, device=src.device) if sc is not None else None
        N.check(N.lib().hl_owner_gather(seg.owner.data_ptr(), src.shape[0], g.data_ptr(), g.stride(0),
                                        seg.inv_count.data_ptr(), N.ptr(sc), src.data_ptr(), src.stride(0),
                                        dsrc.data_ptr(), dsrc.stride(0), N.ptr(dsc), src.shape[1], N.stream_ptr()),
                "hl_owner_gather")
        return dsrc, (None if dsc is None else dsc.view(-1, 1)), None


def segment_mean(src, seg, scale=None):
    """mean over each bucket of (scale[m] * src[m]) -- scatter_mean / global_mean_pool with the gate
    multiply fused (lib/Hodge_ST_Model.py:141-150, :636)."""
    return _SegmentMean.apply(src, scale, seg)


# ---------------------------------------------------------------------------------------------
# BatchNorm (training statistics) + activation, optionally writing into a slice of a wide buffer
# ---------------------------------------------------------------------------------------------
class _BnAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, slope, nvalid, running_mean, running_var, momentum, counter=None, tap=None,
                tiles=None):
        N.require_cuda_f32(x, gamma, beta)
        L = N.lib()
        from_tiles = tiles is not None and tiles.matches(x, nvalid)
        x, ldx = N.row_major(x)
        R, F = x.shape
        # tap = (stack, side, c0, c1): the output is block [c0, c1) of a dense-connection buffer (dense_stack.DenseStack),
        # written in place there, and the backward pass picks the block's gradient up from the stack's accumulator
        ctx.tap = tap
        y = torch.empty((R, F), dtype=torch.float32, device=x.device) if tap is None else tap[0].target(*tap[1:])[0]
        stats = torch.empty(2 * F, dtype=torch.float32, device=x.device)
        if R == 0:                                  # e.g. a coarse level without edges: nothing to normalise
            stats.zero_()
            ctx.empty = True
            ctx.mark_non_differentiable(stats)
            ctx.set_materialize_grads(False)
            return y, stats
        ctx.empty = False
        if from_tiles:                              # statistics already reduced per 32-row block by the producing GEMM
            N.check(L.hl_bn_act_fwd_tiles(x.data_ptr(), ldx, R, F, N.ptr(gamma), N.ptr(beta), eps, slope,
                                          y.data_ptr(), y.stride(0), stats.data_ptr(), N.ptr(nvalid), N.ptr(running_mean),
                                          N.ptr(running_var), float(momentum), N.ptr(counter), tiles.part.data_ptr(),
                                          N.stream_ptr()), "hl_bn_act_fwd_tiles")
        else:
            nb = L.hl_bn_workspace(R, F)
            ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
            N.check(L.hl_bn_act_fwd(x.data_ptr(), ldx, R, F, N.ptr(gamma), N.ptr(beta), eps, slope,
                                    y.data_ptr(), y.stride(0), stats.data_ptr(), N.ptr(nvalid), N.ptr(running_mean),
                                    N.ptr(running_var), float(momentum), N.ptr(counter), ws.data_ptr(), nb, N.stream_ptr()),
                    "hl_bn_act_fwd")
        ctx.eps, ctx.slope, ctx.nvalid = eps, slope, nvalid
        ctx.params = (gamma, beta)
        ctx.save_for_backward(x, y, gamma, stats)
        ctx.mark_non_differentiable(stats)
        ctx.set_materialize_grads(False)           # no zero-filled gradient tensor for `stats` on every backward
        return y, stats

    @staticmethod
    def backward(ctx, dy, _):
        dy2 = None
        if ctx.tap is not None:                     # stack-accumulated gradient + autograd's piece: summed inside the kernels
            dy, dy2 = ctx.tap[0].collect("own", *ctx.tap[1:], dy, fuse=True)
        if dy is None:
            return (None,) * 12
        if ctx.empty:
            return (torch.zeros_like(dy),) + (None,) * 11
        x, y, gamma, stats = ctx.saved_tensors
        L = N.lib()
        R, F = x.shape
        dy, lddy = N.row_major(dy)
        lddy2 = 0
        if dy2 is not None:
            dy2, lddy2 = N.row_major(dy2)
        dx = torch.empty((R, F), dtype=torch.float32, device=x.device)
        tg, tb = _grad_target(ctx.params[0]), _grad_target(ctx.params[1])
        fused = tg is not None and tb is not None
        dgamma = tg if fused else torch.empty(F, dtype=torch.float32, device=x.device)
        dbeta = tb if fused else torch.empty(F, dtype=torch.float32, device=x.device)
        nb = L.hl_bn_workspace(R, F)
        ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
        N.check(L.hl_bn_act_bwd(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), dy.data_ptr(), lddy, N.ptr(dy2), lddy2, R, F,
                                N.ptr(gamma), stats.data_ptr(), ctx.eps, ctx.slope, dx.data_ptr(), dx.stride(0),
                                dgamma.data_ptr(), dbeta.data_ptr(), 1 if fused else 0, N.ptr(ctx.nvalid), ws.data_ptr(), nb,
                                N.stream_ptr()), "hl_bn_act_bwd")
        if fused or ctx.params[0] is None:          # BatchNorm1d(affine=False): gamma / beta are not autograd inputs
            dgamma = None
        if fused or ctx.params[1] is None:
            dbeta = None
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None, None


class _BnActEval(torch.autograd.Function):
    """Inference-mode BatchNorm1d (+ activation): y = act((x - running_mean) rsqrt(running_var + eps) gamma + beta) through
    hl_bn_apply with the running statistics in place of the batch statistics; the statistics are constants, so the
    adjoint is hl_bn_bwd_apply with zero column sums (dx = gamma rstd dz) and hl_bn_bwd_sums for dgamma / dbeta."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, slope, nvalid):
        N.require_cuda_f32(x, gamma, beta, running_mean, running_var)
        L = N.lib()
        x, ldx = N.row_major(x)
        R, F = x.shape
        y = torch.empty((R, F), dtype=torch.float32, device=x.device)
        stats = torch.cat([running_mean.detach().reshape(-1), running_var.detach().reshape(-1)]).contiguous()
        if R > 0:
            N.check(L.hl_bn_apply(x.data_ptr(), ldx, R, F, N.ptr(gamma), N.ptr(beta), stats.data_ptr(), eps, slope,
                                  y.data_ptr(), y.stride(0), N.ptr(nvalid), N.stream_ptr()), "hl_bn_apply")
        ctx.eps, ctx.slope, ctx.nvalid = eps, slope, nvalid
        ctx.has = (gamma is not None, beta is not None)
        ctx.save_for_backward(x, y, gamma, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, stats = ctx.saved_tensors
        L = N.lib()
        R, F = x.shape
        dy, lddy = N.row_major(dy)
        dx = torch.empty((R, F), dtype=torch.float32, device=x.device)
        if R == 0:
            return dx, None, None, None, None, None, None, None
        sums = torch.empty(2 * F, dtype=torch.float32, device=x.device)
        zeros = torch.zeros(2 * F, dtype=torch.float32, device=x.device)
        one = torch.ones(1, dtype=torch.float32, device=x.device)
        nb = L.hl_bn_workspace(R, F)
        ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
        N.check(L.hl_bn_bwd_sums(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), dy.data_ptr(), lddy, R, F,
                                 stats.data_ptr(), ctx.eps, ctx.slope, sums.data_ptr(), N.ptr(ctx.nvalid), ws.data_ptr(), nb,
                                 N.stream_ptr()), "hl_bn_bwd_sums")
        N.check(L.hl_bn_bwd_apply(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), dy.data_ptr(), lddy, R, F,
                                  N.ptr(gamma), stats.data_ptr(), zeros.data_ptr(), one.data_ptr(), ctx.eps, ctx.slope,
                                  dx.data_ptr(), dx.stride(0), N.ptr(ctx.nvalid), N.stream_ptr()), "hl_bn_bwd_apply")
        return dx, (sums[F:] if ctx.has[0] else None), (sums[:F] if ctx.has[1] else None), None, None, None, None, None


def bn_act_eval(x, gamma, beta, running_mean, running_var, eps=1e-5, slope=0.0, nvalid=None):
    """Inference-mode BatchNorm1d over rows + (leaky) ReLU (slope = 1: no activation) with the running statistics."""
    return _BnActEval.apply(x, gamma, beta, running_mean, running_var, float(eps), float(slope), nvalid)


class _SyncBnAct(torch.autograd.Function):
    """Training BatchNorm + (leaky) ReLU with statistics over all ranks of a process group: the kernels of
    `_BnAct` run as separate phases (hl_bn_stats / hl_bn_apply, hl_bn_bwd_sums / hl_bn_bwd_apply) with one small
    exchange between them (parallel.combine_bn_stats / reduce_bn_sums).  dgamma / dbeta are this rank's own
    column sums -- the gradient all-reduce averages them like every other parameter gradient."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, slope, nvalid, group):
        from . import parallel as P
        N.require_cuda_f32(x, gamma, beta)
        L = N.lib()
        x, ldx = N.row_major(x)
        R, F = x.shape
        y = torch.empty((R, F), dtype=torch.float32, device=x.device)
        local = torch.empty(2 * F, dtype=torch.float32, device=x.device)
        nb = L.hl_bn_workspace(R, F)
        ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
        N.check(L.hl_bn_stats(x.data_ptr(), ldx, R, F, local.data_ptr(), N.ptr(nvalid), ws.data_ptr(), nb, N.stream_ptr()),
                "hl_bn_stats")
        count = nvalid.to(torch.float32) if nvalid is not None else torch.full((1,), float(R), device=x.device)
        stats, total = P.combine_bn_stats(local, count, group)
        N.check(L.hl_bn_apply(x.data_ptr(), ldx, R, F, N.ptr(gamma), N.ptr(beta), stats.data_ptr(), eps, slope,
                              y.data_ptr(), y.stride(0), N.ptr(nvalid), N.stream_ptr()), "hl_bn_apply")
        ctx.eps, ctx.slope, ctx.nvalid, ctx.group = eps, slope, nvalid, group
        ctx.params = (gamma, beta)
        inv_total = (1.0 / total.clamp(min=1.0)).reshape(1).contiguous()
        ctx.save_for_backward(x, y, gamma, stats, inv_total)
        ctx.mark_non_differentiable(stats, total)
        ctx.set_materialize_grads(False)
        return y, stats, total

    @staticmethod
    def backward(ctx, dy, _s, _t):
        if dy is None:
            return (None,) * 7
        from . import parallel as P
        x, y, gamma, stats, inv_total = ctx.saved_tensors
        L = N.lib()
        R, F = x.shape
        dy, lddy = N.row_major(dy)
        dx = torch.empty((R, F), dtype=torch.float32, device=x.device)
        sums = torch.empty(2 * F, dtype=torch.float32, device=x.device)
        nb = L.hl_bn_workspace(R, F)
        ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
        N.check(L.hl_bn_bwd_sums(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), dy.data_ptr(), lddy, R, F,
                                 stats.data_ptr(), ctx.eps, ctx.slope, sums.data_ptr(), N.ptr(ctx.nvalid), ws.data_ptr(), nb,
                                 N.stream_ptr()), "hl_bn_bwd_sums")
        gsums = P.reduce_bn_sums(sums, ctx.group)
        N.check(L.hl_bn_bwd_apply(x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), dy.data_ptr(), lddy, R, F,
                                  N.ptr(gamma), stats.data_ptr(), gsums.data_ptr(), inv_total.data_ptr(), ctx.eps, ctx.slope,
                                  dx.data_ptr(), dx.stride(0), N.ptr(ctx.nvalid), N.stream_ptr()), "hl_bn_bwd_apply")
        dbeta, dgamma = sums[:F], sums[F:]
        tg, tb = _grad_target(ctx.params[0]), _grad_target(ctx.params[1])
        if tg is not None and tb is not None:
            tg.add_(dgamma)
            tb.add_(dbeta)
            dgamma = dbeta = None
        if ctx.params[0] is None:
            dgamma = None
        if ctx.params[1] is None:
            dbeta = None
        return dx, dgamma, dbeta, None, None, None, None


def bn_act_train_synced(x, gamma, beta, group, eps=1e-5, slope=0.0, nvalid=None, running_mean=None, running_var=None,
                        momentum=0.1, counter=None):
    """`bn_act_train` with batch statistics over all ranks of `group`; returns (y, stats[2F]).  Running statistics
    follow nn.SyncBatchNorm: global mean, unbiased global variance (N - 1 over all ranks)."""
    y, stats, total = _SyncBnAct.apply(x, gamma, beta, float(eps), float(slope), nvalid, group)
    if running_mean is not None:
        f = running_mean.numel()
        with torch.no_grad():
            unbias = total / (total - 1.0).clamp(min=1.0)
            running_mean.mul_(1.0 - momentum).add_(stats[:f] * momentum)
            running_var.mul_(1.0 - momentum).add_(stats[f:] * (unbias * momentum))
    if counter is not None:
        with torch.no_grad():
            counter += 1
    return y, stats


def bn_act_train(x, gamma, beta, eps=1e-5, slope=0.0, nvalid=None, running_mean=None, running_var=None, momentum=0.1,
                 counter=None, tap=None, tiles=None):
    """Training-mode BatchNorm1d over rows + (leaky) ReLU; returns (y, stats[2F] = mean | biased var).
    `nvalid`: optional device int32 scalar -- rows beyond it are padding (excluded, written as zeros).
    running_mean / running_var (optional) are updated in the same launch, like nn.BatchNorm1d; `counter`
    (optional int64 scalar: num_batches_tracked) is incremented there too.  `tiles` (BnTiles, optional): block statistics
    of x left by the GEMM that produced it (bn_stats_from_epilogue) -- the statistics pass over x is skipped."""
    return _BnAct.apply(x, gamma, beta, float(eps), float(slope), nvalid, running_mean, running_var, momentum, counter, tap, tiles)
