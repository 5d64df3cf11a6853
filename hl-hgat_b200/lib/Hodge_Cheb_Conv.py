"""B200-native operator layer behind the reference's module API (lib/Hodge_Cheb_Conv.py).

Same class names, constructor arguments, attribute / state_dict names and forward signatures
as the reference; the arithmetic runs in libhlhgat.so (sm_100a) -- CUDA tensors only, no
PyG / torch_scatter / torch_sparse, no CPU fallback.
"""
import contextlib
import math

import torch
import torch.nn as nn
from torch import Tensor
from torch.nn import Parameter

from .. import functional as F_hl
from .. import dense_stack as _ds
from .. import lanes as _lanes
from .. import parallel as _parallel
from .. import _native as N
from ..simplex import operator_for, incidence_for, CsrOperator, Incidence

__all__ = ["HodgeLaguerreConv", "HodgeChebConv", "HodgeLaguerreFastConv", "NodeEdgeInt", "MSI",
           "SAPool", "HL_filter", "GraphBatchNorm", "NEConv", "adj2par1", "degree"]


class _GlorotLinear(nn.Module):
    """Stand-in for PyG's dense Linear(bias=False, weight_initializer='glorot'): one `weight`
    [out, in] parameter (state_dict key `lins.{k}.weight`)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = Parameter(torch.empty(out_channels, in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
        with torch.no_grad():
            self.weight.uniform_(-a, a)

    def forward(self, x):
        return F_hl.linear(x.reshape(-1, x.shape[-1]), self.weight).view(*x.shape[:-1], -1)


class _HodgePolyConv(nn.Module):
    family = None

    def __init__(self, in_channels: int, out_channels: int, K: int, bias: bool = True, **kwargs):
        super().__init__()
        assert K > 0
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.lins = nn.ModuleList([_GlorotLinear(in_channels, out_channels) for _ in range(K)])
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        for lin in self.lins:
            lin.reset_parameters()
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, x: Tensor, edge_index, edge_weight=None, batch=None):
        """`edge_index` is the reference's int64 COO [2,nnz] (bucketed once per batch and cached) or
        an already built `CsrOperator`.  `batch` is ignored, as in the reference."""
        op = edge_index if isinstance(edge_index, CsrOperator) else operator_for(edge_index, edge_weight, x.shape[0])
        return F_hl.poly_conv(x, [lin.weight for lin in self.lins], self.bias, op, self.family)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, K={len(self.lins)})"


class HodgeLaguerreConv(_HodgePolyConv):
    """Reference lib/Hodge_Cheb_Conv.py:452-523."""
    family = "laguerre"


class HodgeChebConv(_HodgePolyConv):
    """Reference lib/Hodge_Cheb_Conv.py:366-448.  (The reference's 3-D path raises at :411 and its
    __repr__ raises at :448; here 3-D input is flattened like the Laguerre conv -- A(.) acts on
    columns independently, so the intended result does not depend on the flattening order.)"""
    family = "cheb"


class HodgeLaguerreFastConv(_HodgePolyConv):
    """HL-HGAT-DEMO/lib/Hodge_Cheb_Conv.py:519-582: forward(x, adj_t).  `adj_t` is a CsrOperator or an
    `(edge_index, edge_weight)` pair.  quirk=True reproduces DEMO :561 (every order k >= 2 propagates
    the ORIGINAL x); quirk=False is the HodgeLaguerreConv polynomial."""
    family = "laguerre"

    def __init__(self, in_channels, out_channels, K, bias=True, quirk=True, **kwargs):
        super().__init__(in_channels, out_channels, K, bias)
        self.quirk = quirk

    def forward(self, x, adj_t):
        op = adj_t if isinstance(adj_t, CsrOperator) else operator_for(adj_t[0], adj_t[1], x.shape[0])
        ws = [lin.weight for lin in self.lins]
        if not self.quirk or len(ws) < 3:
            return F_hl.poly_conv(x, ws, self.bias, op, "laguerre")
        return _fastconv_quirk(x, ws, self.bias, op)


def _quirk_coefficients(K):
    """With DEMO :561 every order k >= 2 propagates the ORIGINAL x, so T_k = a_k x + b_k (A x) with
    T_0 = x, T_1 = x - A x, T_{k+1} = (-A x + (2k+1) T_k - k T_{k-1}) / (k+1)."""
    a, b = [1.0, 1.0], [0.0, -1.0]
    for k in range(1, K - 1):
        a.append(((2 * k + 1) * a[k] - k * a[k - 1]) / (k + 1))
        b.append(((2 * k + 1) * b[k] - k * b[k - 1] - 1.0) / (k + 1))
    return a[:K], b[:K]


class _QuirkConv(torch.autograd.Function):
    """out = sum_k T_k W_k^T + b with the quirk basis (DEMO lib/Hodge_Cheb_Conv.py:542-578).  Forward: the basis through
    the SpMM recurrence epilogues in the reference's evaluation order (xg = x for every order), the Theta transforms on
    the tcgen05 GEMM.  Backward: weight gradients g^T T_k on the tensor-core weight-gradient kernel; since T_k = a_k x +
    b_k A x, dx = g (sum_k a_k W_k) + A^T (g (sum_k b_k W_k)): two GEMMs and one adjoint SpMM."""

    @staticmethod
    def forward(ctx, x, bias, op, inner, *weights):
        N.require_cuda_f32(x, bias, *weights)
        K = len(weights)
        x = x.contiguous()                       # [R, T*C]: A acts on rows, the Theta transforms on the last `inner` columns
        ts = [x]
        for k in range(K - 1):
            if k == 0:
                ts.append(F_hl.poly_spmm(op.fwd, op.nrows, x, N.HL_EPI_LAGUERRE_FIRST, p1=x))
            else:
                ts.append(F_hl.poly_spmm(op.fwd, op.nrows, x, N.HL_EPI_LAGUERRE_STEP, c=(float(k), 0, 0, 0),
                                         p1=ts[k], p2=ts[k - 1]))
        v = [t.view(-1, inner) for t in ts]
        out = F_hl.dense2(v[0], weights[0], v[1], weights[1], bias)
        for k in range(2, K, 2):
            if k + 1 < K:
                F_hl.dense2(v[k], weights[k], v[k + 1], weights[k + 1], None, out=out, accumulate=True)
            else:
                F_hl.dense(v[k], weights[k], None, out=out, accumulate=True)
        ctx.op, ctx.has_bias, ctx.inner = op, bias is not None, inner
        ctx.save_for_backward(*ts, *weights)
        return out

    @staticmethod
    def backward(ctx, g):
        K = len(ctx.saved_tensors) // 2
        ts, weights = ctx.saved_tensors[:K], ctx.saved_tensors[K:]
        op, inner = ctx.op, ctx.inner
        g = g.contiguous()
        gb = F_hl.colsum(g) if ctx.has_bias and ctx.needs_input_grad[1] else None
        gws = [F_hl.wgrad(g, ts[k].view(-1, inner)) if ctx.needs_input_grad[4 + k] else None for k in range(K)]
        gx = None
        if ctx.needs_input_grad[0]:
            a, b = _quirk_coefficients(K)
            wa = sum(ak * w for ak, w in zip(a, weights))            # [Fout, Fin] weight-sized combinations
            wb = sum(bk * w for bk, w in zip(b, weights))
            own = F_hl.dense(g, wa.contiguous(), transpose_w=True).view(ts[0].shape)
            acc = F_hl.dense(g, wb.contiguous(), transpose_w=True).view(ts[0].shape)
            gx = F_hl.poly_spmm(op.bwd, op.nrows, acc, N.HL_EPI_LINCOMB, c=(1.0, 1.0, 0, 0), p1=own)
        return (gx, gb, None, None, *gws)


def _fastconv_quirk(x, ws, bias, op):
    shp = x.shape
    out = _QuirkConv.apply(x.reshape(shp[0], -1), bias, op, shp[-1], *ws)
    return out.view(*shp[:-1], -1)


# ---------------------------------------------------------------------------------------------
def adj2par1(edge_index, num_node, num_edge):
    """Reference lib/Hodge_Dataset.py:169-191: B1 as a sparse COO float tensor [N,E], -1 at the tail
    (edge_index[0]), +1 at the head.  The int32 incidence tables the kernels use are built once and
    attached to the returned tensor (`._hl_incidence`)."""
    e = edge_index.shape[1]
    ar = torch.arange(e, device=edge_index.device)
    idx = torch.stack([torch.cat([edge_index[0], edge_index[1]]), torch.cat([ar, ar])])
    val = torch.cat([edge_index.new_full((e,), -1), edge_index.new_full((e,), 1)]).to(torch.float)
    par = torch.sparse_coo_tensor(idx, val, (num_node, num_edge), check_invariants=False)
    if edge_index.is_cuda:
        par._hl_incidence = incidence_for(edge_index, num_node)
    return par


def degree(index, num_nodes=None, dtype=torch.float32):
    """PyG utils.degree (used as `degree(edge_index.view(-1)) [+ 1e-6]`, lib/Hodge_ST_Model.py:624)."""
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    out = torch.zeros(n, dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=dtype, device=index.device))


def _incidence_of(par):
    if isinstance(par, Incidence):
        return par
    inc = getattr(par, "_hl_incidence", None)
    if inc is not None:
        return inc
    # a foreign sparse +-1 matrix: recover (tail, head) per column; cached on the tensor
    pc = par.coalesce()
    idx, val = pc.indices(), pc.values()
    order = torch.argsort(idx[1] * 2 + (val > 0).long(), stable=True)
    rows = idx[0][order].view(-1, 2).t().contiguous()
    inc = incidence_for(rows, par.shape[0])
    par._hl_incidence = inc
    return inc


class NodeEdgeInt(nn.Module):
    """Reference lib/Hodge_Cheb_Conv.py:255-309 (MSI :61-115 is the same module)."""

    def __init__(self, d=64, dk=32, dv=64, dl=64, only_att=False, sigma=nn.Sigmoid(), l=0.9):
        super().__init__()
        dl = dv
        self.sigma = sigma
        self.dk = dk
        self.only_att = only_att
        if only_att:
            self.WQ_Node = nn.Linear(d, dk)
            self.WK_Node = nn.Linear(d, dk)
            self.WQ_Edge = nn.Linear(d, dk)
            self.WK_Edge = nn.Linear(d, dk)
        else:
            self.WV_Node = nn.Sequential(nn.Linear(d * 2, dl), nn.BatchNorm1d(dl), nn.ReLU(),
                                         nn.Linear(dl, dv), nn.BatchNorm1d(dv), nn.ReLU())
            self.WV_Edge = nn.Sequential(nn.Linear(d * 2, dl), nn.BatchNorm1d(dl), nn.ReLU(),
                                         nn.Linear(dl, dv), nn.BatchNorm1d(dv), nn.ReLU())
        self.lambda_Node = l
        self.lambda_Edge = l

    def _sigma_name(self):
        if isinstance(self.sigma, nn.Sigmoid):
            return "sigmoid"
        if isinstance(self.sigma, nn.ReLU):
            return "relu"
        return None

    def forward(self, x_t, x_s, par, D, nvalid=(None, None)):
        inc = _incidence_of(par)
        ln = _lanes.active()
        if not self.only_att:
            s2t = lambda v: F_hl.edge_to_node(v, D, inc)        # (1/D) |B1| v      (:294)  # noqa: E731
            t2s = lambda v: F_hl.node_to_edge(v, inc)           # |B1|^T v / 2      (:295)  # noqa: E731
            if ln is None:
                return _mlp(self.WV_Node, x_s, s2t, x_t, nvalid[0]), _mlp(self.WV_Edge, x_t, t2s, x_s, nvalid[1])
            # two-lane issue (lanes.py): the node MLP stays on the node lane, the edge MLP goes to the edge lane;
            # the transfers are the only place where a lane reads the other lane's features
            ln.exchange(node_tensors=(x_t,), edge_tensors=(x_s,))
            x_t1 = _mlp(self.WV_Node, x_s, s2t, x_t, nvalid[0])
            with ln.edge_ctx():
                x_s1 = _mlp(self.WV_Edge, x_t, t2s, x_s, nvalid[1])
            return x_t1, x_s1
        if ln is not None:
            ln.to_node(x_s)                # the gate is computed on the node lane
        x_s2t = F_hl.edge_to_node(x_s, D, inc)
        x_t2s = F_hl.node_to_edge(x_t, inc)
        proj = lambda lin, v: F_hl.linear(v, lin.weight, lin.bias)        # Q / K projections (:297-302) on the tcgen05 GEMM  # noqa: E731
        k_t, k_s = proj(self.WK_Node, x_t), proj(self.WK_Edge, x_s)
        name = self._sigma_name()
        if name is None:        # arbitrary activation object: gate pre-activation through the kernel is not available
            raise N.HlError("NodeEdgeInt(only_att=True) supports sigma = nn.Sigmoid() or nn.ReLU()")
        a_t = F_hl.att_gate(proj(self.WQ_Edge, x_s2t), proj(self.WQ_Node, x_t), k_t, self.lambda_Node, name)
        a_s = F_hl.att_gate(proj(self.WQ_Node, x_t2s), proj(self.WQ_Edge, x_s), k_s, self.lambda_Edge, name)
        return a_t, a_s


MSI = NodeEdgeInt


def _mlp_stack(seq, stack, side, d, nvalid=None):
    """`_mlp` reading [transferred | own] in place from the dense-connection buffers of `stack`."""
    lin0, bn0, _, lin1, bn1, _ = seq
    with _epilogue_stats(bn0, nvalid) as tiles:
        h = _ds.stack_linear(stack, side, d, lin0.weight, lin0.bias)
    h = _bn_relu(bn0, h, 0.0, nvalid, tiles=tiles)
    with _epilogue_stats(bn1, nvalid) as tiles:
        h = F_hl.linear(h, lin1.weight, lin1.bias)
    return _bn_relu(bn1, h, 0.0, nvalid, tiles=tiles)


def node_edge_int_on_stack(module, stack, nvalid=(None, None)):
    """`NodeEdgeInt.forward(x_t0, x_s0, par, D)` (value path, lib/Hodge_Cheb_Conv.py:293-309) with x_t0 / x_s0 living in
    a `DenseStack`: only the blocks appended since the last call are transferred (each exactly once), the rest of
    `(1/D)|B1| x_s0` / `|B1|^T x_t0 / 2` is already in the stack's transfer buffers."""
    assert not module.only_att
    if isinstance(stack, _ds.CatStack):
        return module(stack.x["t"], stack.x["s"], stack.inc, stack.D, nvalid)
    ln = _lanes.active()
    d = stack.cols
    if ln is not None:
        c0, c1, x_t, x_s = stack.blocks[-1]
        ln.exchange(node_tensors=(x_t,), edge_tensors=(x_s,))
    stack.transfer_pending("t")
    x_t1 = _mlp_stack(module.WV_Node, stack, "t", d, nvalid[0])
    with (ln.edge_ctx() if ln is not None else contextlib.nullcontext()):
        stack.transfer_pending("s")
        x_s1 = _mlp_stack(module.WV_Edge, stack, "s", d, nvalid[1])
    return x_t1, x_s1


def _into(tap, y):
    """A block computed into its own tensor -> its slice of the dense-connection buffer (copied in)."""
    return y if tap is None else _ds._Publish.apply(y, *tap)


def _epilogue_stats(bn, nvalid=None):
    """`with _epilogue_stats(bn, nvalid) as tiles: h = <Linear / conv>` then `_bn_relu(bn, h, ..., tiles=tiles)`: the
    Linear -> BatchNorm1d pairs of the reference (lib/Hodge_Cheb_Conv.py:277-288, lib/Hodge_ST_Model.py:578-601) with
    the batch statistics reduced per 32-row block in the GEMM epilogue instead of by a separate pass over h.  Only for
    the plain training-mode BatchNorm (not eval(), not the synced variant: those do not read batch statistics of h)."""
    if isinstance(bn, GraphBatchNorm):
        bn = bn.module
    on = (bn.training or not bn.track_running_stats) and _parallel.sync_batchnorm_group() is False
    return F_hl.bn_stats_from_epilogue(nvalid, enabled=on)


def _bn_relu(bn, x, slope=0.0, nvalid=None, tap=None, tiles=None):
    """nn.BatchNorm1d (+ReLU) through the fused kernels in training mode; running statistics updated
    exactly like torch (momentum, unbiased variance).  `nvalid` (device int32 scalar) marks the rows
    beyond it as padding of a fixed-capacity batch.  `tap` = (stack, side, c0, c1): the output is written
    straight into that block of a dense-connection buffer (dense_stack.DenseStack).  `tiles`: see _epilogue_stats."""
    if not (bn.training or not bn.track_running_stats):      # eval(): the running statistics, same apply kernel
        return _into(tap, F_hl.bn_act_eval(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, slope, nvalid))
    track = bn.track_running_stats and bn.training
    if track:
        counter = bn.num_batches_tracked
        if bn.momentum is None:               # cumulative average: the factor depends on the counter (host read)
            with torch.no_grad():
                bn.num_batches_tracked += 1
            m, counter = 1.0 / float(bn.num_batches_tracked), None
        else:
            m = bn.momentum                   # the counter is incremented inside the statistics kernel
        rm, rv = bn.running_mean, bn.running_var
    else:
        rm = rv = counter = None
        m = 0.1
    group = _parallel.sync_batchnorm_group()
    if group is not False:                    # opt-in: statistics over all data-parallel ranks (parallel.enable_sync_batchnorm)
        y, _ = F_hl.bn_act_train_synced(x, bn.weight, bn.bias, group, bn.eps, slope, nvalid, rm, rv, m, counter)
        return _into(tap, y)
    y, _ = F_hl.bn_act_train(x, bn.weight, bn.bias, bn.eps, slope, nvalid, rm, rv, m, counter, tap, tiles)
    return y


def _mlp(seq, other, transfer, own, nvalid=None):
    """Linear(cat[transfer(other), own]) -> BN -> ReLU -> Linear -> BN -> ReLU without materialising the
    concat: the first Linear is split over its two column blocks (lib/Hodge_Cheb_Conv.py:307-308).  `other` are the
    features of the other simplex order, `transfer` the linear map that brings them over (:294 / :295).  With
    functional.enable_project_then_transfer() and a layer narrower than `other`, the column block acting on the
    transferred features is applied BEFORE the transfer (it commutes with it: the transfer mixes rows, the weights
    mix columns), so the transfer moves the layer width."""
    lin0, bn0, _, lin1, bn1, _ = seq
    d = other.shape[1]
    if F_hl.project_then_transfer_enabled() and lin0.out_features < d and lin0.out_features % 16 == 0 and d % 4 == 0:
        t = transfer(F_hl.linear_part(other, lin0.weight, 0, d))
        h = F_hl.linear_part(own, lin0.weight, d, lin0.weight.shape[1], lin0.bias, addend=t)
    else:
        t = transfer(other)
        with _epilogue_stats(bn0, nvalid) as tiles:
            h = F_hl.linear(t, lin0.weight, lin0.bias, x2=own)
        h = _bn_relu(bn0, h, 0.0, nvalid, tiles=tiles)
        with _epilogue_stats(bn1, nvalid) as tiles:
            h = F_hl.linear(h, lin1.weight, lin1.bias)
        return _bn_relu(bn1, h, 0.0, nvalid, tiles=tiles)
    h = _bn_relu(bn0, h, 0.0, nvalid)
    with _epilogue_stats(bn1, nvalid) as tiles:
        h = F_hl.linear(h, lin1.weight, lin1.bias)
    return _bn_relu(bn1, h, 0.0, nvalid, tiles=tiles)


class GraphBatchNorm(nn.Module):
    """gnn.BatchNorm: a wrapper whose only child is `.module = nn.BatchNorm1d` (state_dict keys
    `module.weight`, ...).  `forward_act` fuses the following (leaky) ReLU."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum)

    def forward(self, x):
        return _bn_relu(self.module, x, slope=1.0)

    def forward_act(self, x, slope=0.0, nvalid=None, tap=None, tiles=None):
        return _bn_relu(self.module, x, slope, nvalid, tap, tiles)


class NEConv(nn.Module):
    """The reference's 9-entry gnn.Sequential NEConv block (lib/Hodge_ST_Model.py:578-590,
    lib/Hodge_Cheb_Conv.py:142-154): conv_t, BN, act, dropout, conv_s, BN, act, dropout, list-pack.
    Children keep the reference's names module_0 / module_1 / module_4 / module_5."""

    def __init__(self, fin_t, fin_s, fout, K, dropout_ratio=0.0, slope=0.0, conv=HodgeLaguerreConv):
        super().__init__()
        self.module_0 = conv(fin_t, fout, K=K)
        self.module_1 = GraphBatchNorm(fout)
        self.module_4 = conv(fin_s, fout, K=K)
        self.module_5 = GraphBatchNorm(fout)
        self.slope, self.p = slope, dropout_ratio

    def forward(self, x_t, edge_index_t, edge_weight_t, x_s, edge_index_s, edge_weight_s, nvalid=(None, None), stack=None):
        """`stack` (dense_stack.DenseStack, optional): the outputs become the next block of the dense-connection
        buffers -- the BatchNorm kernel writes them there directly (the `torch.cat` of lib/Hodge_ST_Model.py:632-633)."""
        ln = _lanes.active()
        drop = self.p > 0.0 and self.training
        taps = (None, None)
        cat = isinstance(stack, _ds.CatStack)
        if stack is not None and not cat:
            c0, c1 = stack.reserve(self.module_1.module.num_features)
            taps = ((stack, "t", c0, c1), (stack, "s", c0, c1))
        with _epilogue_stats(self.module_1, nvalid[0]) as tiles:
            h_t = self.module_0(x_t, edge_index_t, edge_weight_t)
        x_t = self.module_1.forward_act(h_t, self.slope, nvalid[0], None if drop else taps[0], tiles)
        if self.p > 0.0:
            x_t = torch.nn.functional.dropout(x_t, self.p, self.training)
            if drop:
                x_t = _into(taps[0], x_t)
        with (ln.edge_ctx() if ln is not None else contextlib.nullcontext()):   # x_s lives on the edge lane (lanes.py)
            with _epilogue_stats(self.module_5, nvalid[1]) as tiles:
                h_s = self.module_4(x_s, edge_index_s, edge_weight_s)
            x_s = self.module_5.forward_act(h_s, self.slope, nvalid[1], None if drop else taps[1], tiles)
            if self.p > 0.0:
                x_s = torch.nn.functional.dropout(x_s, self.p, self.training)
                if drop:
                    x_s = _into(taps[1], x_s)
        if cat:
            stack.publish(x_t, x_s)
        elif stack is not None:
            stack.commit(c0, c1, x_t, x_s)
        return [x_t, x_s]


class HL_filter(nn.Module):
    """Reference lib/Hodge_Cheb_Conv.py:117-188."""

    def __init__(self, channels=2, filters=32, K=4, node_dim=64, edge_dim=64, dropout_ratio=0.0,
                 leaky_slope=0.1, if_dense=True):
        super().__init__()
        self.channels, self.filters = channels, filters
        self.node_dim, self.edge_dim, self.if_dense = node_dim, edge_dim, if_dense
        t_in, s_in = node_dim, edge_dim
        for j in range(channels):
            if if_dense:
                setattr(self, f"MSI{j}", MSI(d=t_in, dv=filters))
                setattr(self, f"NEConv{j}", NEConv(filters, filters, filters, K, dropout_ratio, leaky_slope))
                t_in, s_in = t_in + filters, s_in + filters
            else:
                setattr(self, f"NEConv{j}", NEConv(t_in, s_in, filters, K, dropout_ratio, leaky_slope))
                t_in, s_in = filters, filters

    def forward(self, x_t0, edge_index_t, edge_weight_t, x_s0, edge_index_s, edge_weight_s, par_1=None, D=None):
        for j in range(self.channels):
            if self.if_dense:
                x_t, x_s = getattr(self, f"MSI{j}")(x_t0, x_s0, par_1, D)
                x_t, x_s = getattr(self, f"NEConv{j}")(x_t, edge_index_t, edge_weight_t, x_s, edge_index_s, edge_weight_s)
                x_t0 = torch.cat([x_t0, x_t], dim=-1)
                x_s0 = torch.cat([x_s0, x_s], dim=-1)
            else:
                x_t0, x_s0 = getattr(self, f"NEConv{j}")(x_t0, edge_index_t, edge_weight_t, x_s0, edge_index_s, edge_weight_s)
        return x_t0, x_s0


class SAPool(nn.Module):
    """Reference lib/Hodge_Cheb_Conv.py:36-59: gate -> cluster mean (edges with +inf id dropped) ->
    switch to the level-(k+1) operators."""

    def __init__(self, d=64, dk=32):
        super().__init__()
        self.NEAtt = MSI(d=d, dk=dk, only_att=True, sigma=nn.Sigmoid())

    def forward(self, x_t0, x_s0, par_1, D, datas, pos_ts, pos_ss, k, device="cuda:0"):
        att_t, att_s = self.NEAtt(x_t0, x_s0, par_1, D)
        x_t0 = F_hl.segment_mean(x_t0, F_hl.Segments.from_index(pos_ts[k]), att_t)
        x_s0 = F_hl.segment_mean(x_s0, F_hl.Segments.from_index(pos_ss[k]), att_s)
        nxt = datas[k + 1]
        edge_index_s, edge_weight_s = nxt.edge_index_s.to(device), nxt.edge_weight_s.to(device)
        edge_index_t, edge_weight_t = nxt.edge_index_t.to(device), nxt.edge_weight_t.to(device)
        k += 1
        ei = datas[k].edge_index.to(device)
        par_1 = adj2par1(ei, x_t0.shape[0], x_s0.shape[0])
        D = degree(ei.view(-1), num_nodes=x_t0.shape[0]) + 1e-6
        return x_t0, x_s0, par_1, D, k, edge_index_t, edge_weight_t, edge_index_s, edge_weight_s, att_t, att_s
