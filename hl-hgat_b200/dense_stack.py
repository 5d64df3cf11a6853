"""Dense-connection buffers of one coarsening level: no `torch.cat`, every block transferred once.

The reference grows `x_t0 = cat([x_t0, x_t])` / `x_s0 = cat([x_s0, x_s])` after every layer and every `NodeEdgeInt`
re-transfers the WHOLE running concat to the other simplex order (lib/Hodge_ST_Model.py:629-633,
lib/Hodge_Cheb_Conv.py:294-295): 64 + 128 + 192 + 320 + 448 + 704 = 1,856 columns per side for the ZINC stack, of which
only 704 are new.  The transfers act on columns independently, so here

  * the BatchNorm kernel of every layer writes its output block straight into a column slice of ONE preallocated
    `[rows, width]` buffer per side (`own`), and the first Linear of the next `NodeEdgeInt` reads the leading `d` columns
    of it in place (a strided GEMM operand) -- the concat never happens;
  * each block is transferred once into the matching slice of a second buffer (`tr`: `(1/D)|B1| x_s` on the node side,
    `|B1|^T x_t / 2` on the edge side), bit-identical to the columns the reference recomputes;
  * in the backward pass the data gradients of all consumers of a block are accumulated by the GEMM epilogues into one
    gradient buffer per side (`+=` in the tensor-core kernel, no autograd `add` kernels), and every transferred block is
    sent back through the adjoint transfer ONCE, with the summed gradient.

Autograd sees the blocks as ordinary tensors (storage aliases of the buffers, created with `Tensor.set_`, so no view
tracking) and the consumers list them as inputs, which gives the engine the right execution ORDER; the gradient VALUES
travel through the side buffers, and the producer of a block (`functional._BnAct` with a `tap`, or `_Publish`) picks its
gradient up from there.  All writes go through the raw kernels, so version counters never move.
"""
import torch

from . import _native as N
from . import functional as F_hl
from . import lanes as _lanes

_ENABLED = {"on": True}


def enable_dense_stack(flag=True):
    """On (default): the mirrored model classes keep their dense connections in `DenseStack` buffers.  Off: the
    reference's `torch.cat` + full re-transfer per layer (for A/B runs; forward results are bit-identical)."""
    _ENABLED["on"] = bool(flag)


def dense_stack_enabled():
    return _ENABLED["on"]


def _alias(buf, c0, c1):
    """Columns [c0, c1) of `buf` as a fresh tensor sharing its storage (not an autograd view)."""
    return torch.empty(0, dtype=buf.dtype, device=buf.device).set_(buf.untyped_storage(), buf.storage_offset() + c0,
                                                                   (buf.shape[0], c1 - c0), (buf.stride(0), 1))


class DenseStack:
    def __init__(self, n_t, n_s, width, inc, D, device):
        self.W = (int(width) + 3) // 4 * 4                  # 16-byte aligned rows (TMA operands, 128-bit loads)
        self.rows = {"t": int(n_t), "s": int(n_s)}
        self.inc, self.D, self.device = inc, D, device
        mk = lambda r: torch.empty((r, self.W), dtype=torch.float32, device=device)      # noqa: E731
        self.own = {"t": mk(n_t), "s": mk(n_s)}             # the dense-connection buffers x_t0 / x_s0
        self.tr = {"t": mk(n_t), "s": mk(n_s)}              # tr["t"] = (1/D)|B1| x_s0 (:294), tr["s"] = |B1|^T x_t0 / 2 (:295)
        self.g_own = {"t": None, "s": None}                 # gradient accumulators (allocated by the backward pass)
        self.g_tr = {"t": None, "s": None}
        self.filled = {("own", "t"): 0, ("own", "s"): 0, ("tr", "t"): 0, ("tr", "s"): 0}
        self.consumed = {("own", "t"): 0, ("own", "s"): 0, ("tr", "t"): 0, ("tr", "s"): 0}
        self.cols = 0                                       # columns published (the same on both sides)
        self.blocks = []                                    # (c0, c1, x_t block, x_s block): autograd tensors
        self.tr_blocks = {"t": [], "s": []}                 # transferred blocks (autograd tensors, for ordering)
        ln = _lanes.active()
        cur = torch.cuda.current_stream(device)
        self.stream = {"t": cur, "s": ln.edge if ln is not None else cur}
        if ln is not None:
            for b in (*self.own.values(), *self.tr.values()):
                b.record_stream(ln.edge)

    # ---- forward side ------------------------------------------------------------------------------------------
    def reserve(self, width):
        c0 = self.cols
        if c0 + width > self.W:
            raise N.HlError(f"DenseStack overflow: {c0} + {width} > {self.W} columns")
        return c0, c0 + width

    def target(self, side, c0, c1):
        """Where the producing kernel writes block [c0, c1) of `side`, and the tap its backward reads the gradient from."""
        return _alias(self.own[side], c0, c1), (self, side, c0, c1)

    def commit(self, c0, c1, x_t, x_s):
        assert c0 == self.cols
        self.blocks.append((c0, c1, x_t, x_s))
        self.cols = c1

    def publish(self, x_t, x_s):
        """Append blocks computed elsewhere (copied in)."""
        c0, c1 = self.reserve(x_t.shape[1])
        a_t = _Publish.apply(x_t, self, "t", c0, c1)
        lane = self.stream["s"]
        cur = torch.cuda.current_stream(self.device)
        if lane != cur:                                      # x_s may have been produced on the node lane (pooling, gates)
            lane.wait_stream(cur)
            x_s.record_stream(lane)
        with torch.cuda.stream(lane):
            a_s = _Publish.apply(x_s, self, "s", c0, c1)
        self.commit(c0, c1, a_t, a_s)
        return a_t, a_s

    def view(self, kind, side, c0, c1):
        return _alias((self.own if kind == "own" else self.tr)[side], c0, c1)

    def transfer_pending(self, side):
        """Transfer every block of the OTHER side that has not been brought over to `side` yet."""
        done = len(self.tr_blocks[side])
        for c0, c1, x_t, x_s in self.blocks[done:]:
            self.tr_blocks[side].append(_StackTransfer.apply(x_s if side == "t" else x_t, self, side, c0, c1))

    def deps(self, side):
        own = [b[2] if side == "t" else b[3] for b in self.blocks]
        return own + list(self.tr_blocks[side])

    def close(self):
        """End of the forward pass over this stack: drop the references to the block tensors.  They were only held to
        name them as inputs of later consumers; keeping them would close a reference cycle (block -> grad_fn -> ctx ->
        stack -> block) that keeps the whole autograd graph of a step -- and the AccumulateGrad nodes of the parameters,
        with the stream they were created on -- alive until the garbage collector runs."""
        self.blocks = []
        self.tr_blocks = {"t": [], "s": []}

    def whole(self, side):
        """The dense-connection buffer of `side` as one autograd tensor [rows, cols] (for gates, pooling, heads)."""
        return _StackView.apply(self, side, self.cols, *[b[2] if side == "t" else b[3] for b in self.blocks])

    # ---- backward side -----------------------------------------------------------------------------------------
    def _gbuf(self, kind, side):
        store = self.g_own if kind == "own" else self.g_tr
        if store[side] is None:
            store[side] = torch.empty((self.rows[side], self.W), dtype=torch.float32, device=self.device)
            for st in self.stream.values():
                store[side].record_stream(st)
        return store[side]

    def accumulate_dgrad(self, kind, side, d, g, w):
        """g_buf[:, :d] (=|+=) g @ w   (w: [F, d] column block of the consuming Linear's weight)."""
        gbuf = self._gbuf(kind, side)
        f = min(self.filled[(kind, side)], d)
        if f > 0:
            F_hl.dense(g, w[:, :f], out=_alias(gbuf, 0, f), accumulate=True, transpose_w=True)
        if f < d:
            F_hl.dense(g, w[:, f:d], out=_alias(gbuf, f, d), accumulate=False, transpose_w=True)
        self.filled[(kind, side)] = max(self.filled[(kind, side)], d)

    def accumulate_tensor(self, kind, side, d, g):
        gbuf = self._gbuf(kind, side)
        f = min(self.filled[(kind, side)], d)
        if f > 0:
            _alias(gbuf, 0, f).add_(g[:, :f])
        if f < d:
            _alias(gbuf, f, d).copy_(g[:, f:d])
        self.filled[(kind, side)] = max(self.filled[(kind, side)], d)

    def collect(self, kind, side, c0, c1, g, fuse=False):
        """Total gradient of block [c0, c1): what the stack consumers accumulated (if any) plus autograd's `g`.
        fuse=True: returns the two pieces `(a, b)` (b may be None) for a consumer that adds them itself."""
        # `filled`: leading columns some consumer's backward has written.  A consumer whose output did not reach the loss
        # (e.g. the gate of the last stage when only the prediction is trained) never runs and leaves no gradient here;
        # the autograd engine runs every consumer that does BEFORE the producer of the block (they are its graph children)
        if self.filled[(kind, side)] < c1:
            return (g, None) if fuse else g
        sl = _alias(self._gbuf(kind, side), c0, c1)
        if fuse:
            return sl, g
        if g is not None:
            sl.add_(g)
        return sl


class CatStack:
    """The reference's scheme (`torch.cat` per layer, the whole concat re-transferred by every NodeEdgeInt) behind the
    DenseStack interface, for A/B runs (`enable_dense_stack(False)`)."""

    def __init__(self, inc, D):
        self.inc, self.D = inc, D
        self.x = {"t": None, "s": None}
        self.cols = 0

    def publish(self, x_t, x_s):
        ln = _lanes.active()
        self.x["t"] = x_t if self.x["t"] is None else torch.cat([self.x["t"], x_t], dim=-1)
        with (ln.edge_ctx() if ln is not None else torch.cuda.stream(torch.cuda.current_stream())):
            self.x["s"] = x_s if self.x["s"] is None else torch.cat([self.x["s"], x_s], dim=-1)
        self.cols = self.x["t"].shape[1]
        return x_t, x_s

    def whole(self, side):
        return self.x[side]

    def close(self):
        pass


def new_stack(n_t, n_s, width, inc, D, device):
    return DenseStack(n_t, n_s, width, inc, D, device) if _ENABLED["on"] else CatStack(inc, D)


class _Publish(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, stack, side, c0, c1):
        N.require_cuda_f32(x)
        out = _alias(stack.own[side], c0, c1)
        out.copy_(x)
        ctx.stack, ctx.where = stack, (side, c0, c1)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, g):
        g = ctx.stack.collect("own", *ctx.where, g)
        return g, None, None, None, None


class _StackTransfer(torch.autograd.Function):
    """One block to the other simplex order, written into its slice of the `tr` buffer (lib/Hodge_Cheb_Conv.py:294-295)."""

    @staticmethod
    def forward(ctx, x, stack, side, c0, c1):
        inc = stack.inc
        out = _alias(stack.tr[side], c0, c1)
        if side == "t":                                      # (1/D) |B1| x_s
            F_hl._segment_reduce(inc.rowptr, inc.edge, inc.num_nodes, x, N.HL_POST_RCP_ROW, row_scale=stack.D, out=out)
        else:                                                # |B1|^T x_t / 2
            F_hl._endpoint_gather(inc, x, None, 0.5, out=out)
        ctx.stack, ctx.where = stack, (side, c0, c1)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, g):
        stack = ctx.stack
        side = ctx.where[0]
        g = stack.collect("tr", *ctx.where, g)
        if g is None:
            return None, None, None, None, None
        inc = stack.inc
        if side == "t":                                      # adjoint of (1/D)|B1|: gather the two endpoints, scaled by 1/D
            gx = F_hl._endpoint_gather(inc, g, stack.D, 1.0)
        else:
            gx = F_hl._segment_reduce(inc.rowptr, inc.edge, inc.num_nodes, g, N.HL_POST_CONST, cscale=0.5)
        return gx, None, None, None, None


class _StackLinear(torch.autograd.Function):
    """h = [tr[:, :d] | own[:, :d]] W^T + b: the first Linear of a NodeEdgeInt MLP (lib/Hodge_Cheb_Conv.py:307-308) reading
    both halves of its input in place from the stack buffers; the data gradients go to the gradient buffers."""

    @staticmethod
    def forward(ctx, stack, side, d, weight, bias, *deps):
        N.require_cuda_f32(weight, bias)
        xa, xb = stack.view("tr", side, 0, d), stack.view("own", side, 0, d)
        y = F_hl.dense2(xa, weight[:, :d], xb, weight[:, d:], bias, bn=F_hl._take_bn_request())
        for kind in ("tr", "own"):
            stack.consumed[(kind, side)] = max(stack.consumed[(kind, side)], d)
        ctx.stack, ctx.side, ctx.d, ctx.has_bias, ctx.ndeps = stack, side, d, bias is not None, len(deps)
        ctx.params = (weight, bias)
        ctx.save_for_backward(weight)
        return y

    @staticmethod
    def backward(ctx, g):
        (weight,) = ctx.saved_tensors
        stack, side, d = ctx.stack, ctx.side, ctx.d
        g = g.contiguous()
        xa, xb = stack.view("tr", side, 0, d), stack.view("own", side, 0, d)
        gw = gbias = None
        want_bias = ctx.has_bias and ctx.needs_input_grad[4]
        tgt_b = F_hl._grad_target(ctx.params[1]) if want_bias else None
        if ctx.needs_input_grad[3]:
            tgt = F_hl._grad_target(ctx.params[0])
            fold = want_bias and (tgt is None) == (tgt_b is None)
            if fold:
                gbias = tgt_b if tgt_b is not None else torch.empty(weight.shape[0], dtype=torch.float32, device=g.device)
            with F_hl._wgrad_lane(tgt is not None, g, xa, xb):
                gw = torch.empty_like(weight) if tgt is None else tgt
                F_hl.wgrad2(g, xa, xb, gw[:, :d], gw[:, d:], accumulate=tgt is not None, bias_out=gbias if fold else None,
                            bias_accumulate=tgt_b is not None)
            if tgt is not None:
                gw = None
            if fold:
                want_bias = False
                if tgt_b is not None:
                    gbias = None
        if want_bias:
            with F_hl._wgrad_lane(tgt_b is not None, g):
                gbias = F_hl.colsum(g, out=tgt_b)
            if tgt_b is not None:
                gbias = None
        stack.accumulate_dgrad("tr", side, d, g, weight[:, :d])
        stack.accumulate_dgrad("own", side, d, g, weight[:, d:])
        return (None, None, None, gw, gbias) + (None,) * ctx.ndeps


class _StackView(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stack, side, d, *deps):
        stack.consumed[("own", side)] = max(stack.consumed[("own", side)], d)
        ctx.stack, ctx.side, ctx.d, ctx.ndeps = stack, side, d, len(deps)
        return stack.view("own", side, 0, d)

    @staticmethod
    def backward(ctx, g):
        stack, side = ctx.stack, ctx.side
        # the gradient buffer of a side belongs to that side's lane: later `+=` of the lane's GEMMs must be ordered after this
        lane = stack.stream[side]
        cur = torch.cuda.current_stream(g.device)
        if lane != cur:
            lane.wait_stream(cur)
            g.record_stream(lane)
        with torch.cuda.stream(lane):
            stack.accumulate_tensor("own", side, ctx.d, g)
        return (None, None, None) + (None,) * ctx.ndeps


def stack_linear(stack, side, d, weight, bias):
    return _StackLinear.apply(stack, side, d, weight, bias, *stack.deps(side))
