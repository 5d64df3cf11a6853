"""CPU oracle for the HL-HGAT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-torch (CPU, fp32 or fp64) restatement of what the reference computes on the path
named in BASELINE.json: the Hodge-Laplacian polynomial convolutions, the node<->edge simplex
transfer, the attention gate / cluster pooling, the readout and the simplex-graph
construction.  Every function cites the reference lines it follows (paths relative to
/root/reference).  Nothing under `hl-hgat_b200/` may import this file; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs do.

Parity pinning: the reference ships no tests.  This restatement is pinned
against (i) outputs of the UNMODIFIED reference modules imported in the build container through
`oracle/pyg_shim` (committed as `tests/golden/*.pt` by `tests/golden/make_golden.py`), and
(ii) the analytic known answers of SURVEY.md section 4.  The third-party arithmetic
(torch-geometric / torch-scatter 2.0.9 / torch-sparse 0.6.15) is restated from its published
semantics, so parity versus a real PyG install remains *unpinned* -- see DESIGN.md.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn

INF = float("inf")


# --------------------------------------------------------------------------------------
# third-party primitive semantics
# --------------------------------------------------------------------------------------
def propagate(x, edge_index, norm):
    """PyG MessagePassing(aggr='add').propagate with message = norm * x_j
    (lib/Hodge_Cheb_Conv.py:494,502,518-519): gather rows edge_index[0], scale (the product
    is materialised in the working dtype), sum into rows edge_index[1]."""
    msg = norm.view(-1, 1) * x.index_select(0, edge_index[0])
    return torch.zeros_like(x).index_add_(0, edge_index[1], msg)


def degree(index, num_nodes=None, dtype=torch.float32):
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    out = torch.zeros(n, dtype=dtype)
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=dtype))


def scatter_mean(src, index, dim_size=None):
    """torch_scatter.scatter_mean(src[R,F], index[R] or [R,1], dim=0)."""
    index = index.view(-1).long()
    n = (int(index.max()) + 1 if index.numel() else 0) if dim_size is None else dim_size
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype).index_add_(0, index, src)
    cnt = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones(index.numel(), dtype=src.dtype))
    cnt = cnt.clamp(min=1)
    return out / cnt.view((-1,) + (1,) * (src.dim() - 1))


def global_mean_pool(x, batch, size=None):
    return scatter_mean(x, batch, size)


def dense_to_sparse(a):
    idx = a.nonzero().t().contiguous()
    return idx, a[idx[0], idx[1]]


def to_undirected_min(edge_index, edge_attr=None, num_nodes=None):
    """to_undirected(..., reduce='min') (lib/Hodge_Dataset.py:447): both directions,
    sorted by row*N+col, duplicate attrs min-reduced."""
    n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)
    row = torch.cat([edge_index[0], edge_index[1]])
    col = torch.cat([edge_index[1], edge_index[0]])
    key = row * n + col
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    ei = torch.stack([uniq // n, uniq % n])
    if edge_attr is None:
        return ei, None
    ea = torch.cat([edge_attr, edge_attr], 0)
    out = torch.empty((uniq.numel(),) + tuple(ea.shape[1:]), dtype=ea.dtype)
    idx = inv.view((-1,) + (1,) * (ea.dim() - 1)).expand_as(ea)
    out = out.scatter_reduce_(0, idx, ea, "amin", include_self=False)
    return ei, out


# --------------------------------------------------------------------------------------
# boundary operator, construction, collation
# --------------------------------------------------------------------------------------
def adj2par1(edge_index, num_node, num_edge, dtype=torch.float32):
    """lib/Hodge_Dataset.py:169-191: B1 as an (uncoalesced) sparse COO [N,E]; entry
    (edge_index[0][e], e) = -1 (tail) and (edge_index[1][e], e) = +1 (head)."""
    e = edge_index.shape[1]
    ar = torch.arange(e)
    idx = torch.stack([torch.cat([edge_index[0], edge_index[1]]), torch.cat([ar, ar])])
    val = torch.cat([-torch.ones(e, dtype=dtype), torch.ones(e, dtype=dtype)])
    return torch.sparse_coo_tensor(idx, val, (num_node, num_edge))


def build_simplex_graph(edge_index_directed, num_nodes, dtype=torch.float32):
    """Per-graph simplex-graph construction, lib/Hodge_Dataset.py:447-456,467-468
    (same tail in MLGC :276-288): undirected i<j edge list in lexicographic order, dense B1,
    L0 = B1 B1^T, lambda_max by dense eigh, 2 L/lambda_max, row-major COO of the nonzeros."""
    ei, _ = to_undirected_min(edge_index_directed, None, num_nodes)
    ei = ei[:, ei[0] < ei[1]]
    par1 = adj2par1(ei, num_nodes, ei.shape[1], dtype).to_dense()
    l0 = par1 @ par1.T
    maxeig = torch.linalg.eigh(l0)[0].max()
    l0 = 2 * (par1 @ par1.T) / maxeig
    l1 = 2 * (par1.T @ par1) / maxeig
    ei_t, ew_t = dense_to_sparse(l0)
    ei_s, ew_s = dense_to_sparse(l1)
    return SimpleNamespace(edge_index=ei, edge_index_t=ei_t, edge_weight_t=ew_t,
                           edge_index_s=ei_s, edge_weight_s=ew_s, maxeig=maxeig,
                           num_node1=num_nodes, num_edge1=ei.shape[1])


def collate(graphs):
    """PairData.__inc__ + PyG collate (lib/Hodge_Dataset.py:40-48): block-diagonal batch.
    `graphs` = list of namespaces with x_t, x_s, edge_index, edge_index_t/_s, edge_weight_t/_s,
    y.  edge_index/_t shift by #nodes, edge_index_s by #edges."""
    n_off = e_off = 0
    cat = {k: [] for k in ("x_t", "x_s", "edge_index", "edge_index_t", "edge_index_s",
                           "edge_weight_t", "edge_weight_s", "y")}
    nn1, ne1 = [], []
    for g in graphs:
        cat["x_t"].append(g.x_t)
        cat["x_s"].append(g.x_s)
        cat["edge_index"].append(g.edge_index + n_off)
        cat["edge_index_t"].append(g.edge_index_t + n_off)
        cat["edge_index_s"].append(g.edge_index_s + e_off)
        cat["edge_weight_t"].append(g.edge_weight_t)
        cat["edge_weight_s"].append(g.edge_weight_s)
        if getattr(g, "y", None) is not None:
            cat["y"].append(g.y.view(1, -1) if g.y.dim() < 2 else g.y)
        nn1.append(g.x_t.shape[0])
        ne1.append(g.x_s.shape[0])
        n_off += g.x_t.shape[0]
        e_off += g.x_s.shape[0]
    b = SimpleNamespace()
    for k, v in cat.items():
        if not v:
            setattr(b, k, None)
        elif "index" in k:
            setattr(b, k, torch.cat(v, dim=-1))
        else:
            setattr(b, k, torch.cat(v, dim=0))
    b.num_node1 = torch.tensor(nn1)
    b.num_edge1 = torch.tensor(ne1)
    b.num_graphs = len(graphs)
    return b


# --------------------------------------------------------------------------------------
# polynomial Hodge-Laplacian convolutions
# --------------------------------------------------------------------------------------
class _GlorotLinear(nn.Module):
    """PyG dense Linear(bias=False, weight_initializer='glorot'): weight [out,in]."""

    def __init__(self, fin, fout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fout, fin))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, x):
        return x @ self.weight.t()


class _PolyConv(nn.Module):
    def __init__(self, in_channels, out_channels, K, bias=True):
        super().__init__()
        assert K > 0
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lins = nn.ModuleList([_GlorotLinear(in_channels, out_channels) for _ in range(K)])
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)


class HodgeLaguerreConv(_PolyConv):
    """lib/Hodge_Cheb_Conv.py:480-515.  T0 = x; T1 = x - A x;
    T_{k+1} = (-A T_k + (2k+1) T_k - k T_{k-1})/(k+1); out = sum_k T_k W_k^T + b.
    3-D input [N,T,C] is flattened to [N,T*C] for A(.) (:493-496,:500-505)."""

    def forward(self, x, edge_index, edge_weight=None, batch=None):
        shp = x.shape
        t0 = x
        out = self.lins[0](t0)
        t1 = x
        k = 1
        if len(self.lins) > 1:
            xf = x.reshape(shp[0], -1)
            t1 = (xf - propagate(xf, edge_index, edge_weight)).view(shp)
            out = out + self.lins[1](t1)
        for lin in self.lins[2:]:
            a = propagate(t1.reshape(shp[0], -1), edge_index, edge_weight).view(shp)
            t2 = (-a + (2 * k + 1) * t1 - k * t0) / (k + 1)
            k += 1
            out = out + lin(t2)
            t0, t1 = t1, t2
        if self.bias is not None:
            out = out + self.bias
        return out


class HodgeChebConv(_PolyConv):
    """lib/Hodge_Cheb_Conv.py:394-439.  T1 = A x; T_k = 2 A T_{k-1} - T_{k-2} (no L - I
    rescale).  The 3-D path transposes to [N,C,T] before flattening (:409-414,:421-428),
    which is a pure permutation of the columns A(.) acts on independently."""

    def forward(self, x, edge_index, edge_weight=None, batch=None):
        shp = x.shape
        t0 = x
        out = self.lins[0](t0)
        t1 = x
        if len(self.lins) > 1:
            t1 = self._prop(x, edge_index, edge_weight)
            out = out + self.lins[1](t1)
        for lin in self.lins[2:]:
            t2 = 2.0 * self._prop(t1, edge_index, edge_weight) - t0
            out = out + lin(t2)
            t0, t1 = t1, t2
        if self.bias is not None:
            out = out + self.bias
        return out

    @staticmethod
    def _prop(v, edge_index, edge_weight):
        if v.dim() == 3:
            n, t, c = v.shape
            f = v.transpose(1, 2).reshape(n, -1)
            return propagate(f, edge_index, edge_weight).view(n, c, t).transpose(1, 2)
        return propagate(v, edge_index, edge_weight)


class HodgeLaguerreFastConv(_PolyConv):
    """HL-HGAT-DEMO/lib/Hodge_Cheb_Conv.py:542-578: same polynomial through a CSR SpMM on
    `adj_t` (here: (edge_index, edge_weight) of the operator).  Line :561 propagates the
    ORIGINAL x for every k >= 2 (not T_k); `quirk=True` (default) reproduces that."""

    def __init__(self, in_channels, out_channels, K, bias=True, quirk=True):
        super().__init__(in_channels, out_channels, K, bias)
        self.quirk = quirk

    def forward(self, x, edge_index, edge_weight):
        shp = x.shape
        xf = x.reshape(shp[0], -1)
        t0 = x
        out = self.lins[0](t0)
        t1 = x
        k = 1
        if len(self.lins) > 1:
            t1 = (xf - propagate(xf, edge_index, edge_weight)).view(shp)
            out = out + self.lins[1](t1)
        for lin in self.lins[2:]:
            src = xf if self.quirk else t1.reshape(shp[0], -1)
            a = propagate(src, edge_index, edge_weight).view(shp)
            t2 = (-a + (2 * k + 1) * t1 - k * t0) / (k + 1)
            k += 1
            out = out + lin(t2)
            t0, t1 = t1, t2
        if self.bias is not None:
            out = out + self.bias
        return out


# --------------------------------------------------------------------------------------
# node <-> edge simplex transfer, value MLP and attention gate
# --------------------------------------------------------------------------------------
def transfer(x_t, x_s, par, D):
    """lib/Hodge_Cheb_Conv.py:294-295: x_s2t = (1/D) * (|B1| x_s); x_t2s = (|B1|^T x_t)/2."""
    pa = par.abs()
    x_s2t = (1 / D).view(-1, 1) * torch.sparse.mm(pa, x_s)
    x_t2s = torch.sparse.mm(pa.transpose(0, 1), x_t) / 2
    return x_s2t, x_t2s


class NodeEdgeInt(nn.Module):
    """lib/Hodge_Cheb_Conv.py:255-309 (MSI :61-115 is the same module)."""

    def __init__(self, d=64, dk=32, dv=64, dl=64, only_att=False, sigma=None, l=0.9):
        super().__init__()
        dl = dv
        self.sigma = nn.Sigmoid() if sigma is None else sigma
        self.dk, self.only_att = dk, only_att
        if only_att:
            self.WQ_Node, self.WK_Node = nn.Linear(d, dk), nn.Linear(d, dk)
            self.WQ_Edge, self.WK_Edge = nn.Linear(d, dk), nn.Linear(d, dk)
        else:
            def mlp():
                return nn.Sequential(nn.Linear(2 * d, dl), nn.BatchNorm1d(dl), nn.ReLU(),
                                     nn.Linear(dl, dv), nn.BatchNorm1d(dv), nn.ReLU())
            self.WV_Node, self.WV_Edge = mlp(), mlp()
        self.lambda_Node = self.lambda_Edge = l

    def forward(self, x_t, x_s, par, D):
        x_s2t, x_t2s = transfer(x_t, x_s, par, D)
        if self.only_att:
            ln, le, rt = self.lambda_Node, self.lambda_Edge, math.sqrt(self.dk)
            kt, ks = self.WK_Node(x_t), self.WK_Edge(x_s)
            a_t = self.sigma(((1 - ln) * (self.WQ_Edge(x_s2t) * kt).sum(1, keepdim=True)
                              + ln * (self.WQ_Node(x_t) * kt).sum(1, keepdim=True)) / rt)
            a_s = self.sigma(((1 - le) * (self.WQ_Node(x_t2s) * ks).sum(1, keepdim=True)
                              + le * (self.WQ_Edge(x_s) * ks).sum(1, keepdim=True)) / rt)
            return a_t, a_s
        return (self.WV_Node(torch.cat([x_s2t, x_t], -1)),
                self.WV_Edge(torch.cat([x_t2s, x_s], -1)))


MSI = NodeEdgeInt


def attention_pool(x_t0, x_s0, att_t, att_s, pos_t, pos_s):
    """The pooling block lib/Hodge_ST_Model.py:141-150 / lib/Hodge_Cheb_Conv.py:47-53:
    gate, cluster mean of nodes, drop edges whose float cluster id is inf, cluster mean."""
    x_t0 = scatter_mean(x_t0 * att_t, pos_t.to(torch.long))
    keep = ~torch.isinf(pos_s).view(-1)
    x_s0 = scatter_mean((x_s0 * att_s)[keep], pos_s[keep].to(torch.long))
    return x_t0, x_s0


# --------------------------------------------------------------------------------------
# containers reproducing the reference's state_dict names
# --------------------------------------------------------------------------------------
class GraphBatchNorm(nn.Module):
    """gnn.BatchNorm: single child `.module = nn.BatchNorm1d`."""

    def __init__(self, c):
        super().__init__()
        self.module = nn.BatchNorm1d(c)

    def forward(self, x):
        return self.module(x)


class NEConvBlock(nn.Module):
    """The 9-entry gnn.Sequential of lib/Hodge_ST_Model.py:578-590: children named by their
    list position -- module_0 (conv_t), module_1 (BN), module_4 (conv_s), module_5 (BN)."""

    def __init__(self, fin_t, fin_s, fout, K, dropout=0.0, conv=HodgeLaguerreConv, act=None):
        super().__init__()
        self.module_0 = conv(fin_t, fout, K)
        self.module_1 = GraphBatchNorm(fout)
        self.module_4 = conv(fin_s, fout, K)
        self.module_5 = GraphBatchNorm(fout)
        self.act = nn.ReLU() if act is None else act
        self.drop = nn.Dropout(dropout)

    def forward(self, x_t, ei_t, ew_t, x_s, ei_s, ew_s):
        x_t = self.drop(self.act(self.module_1(self.module_0(x_t, ei_t, ew_t))))
        x_s = self.drop(self.act(self.module_5(self.module_4(x_s, ei_s, ew_s))))
        return x_t, x_s


class HL_HGCNN_zinc_dense_int3_pyr(nn.Module):
    """lib/Hodge_ST_Model.py:544-646 (forward :608-646).  `data` needs x_t, x_s,
    edge_index, edge_index_t/_s, edge_weight_t/_s, num_node1, num_edge1."""

    def __init__(self, channels=(2, 2, 2, 2), filters=(64, 128, 256, 512), mlp_channels=(), K=2,
                 node_dim=21, edge_dim=3, num_classes=1, dropout_ratio=0.0,
                 dropout_ratio_mlp=0.0, keig=7):
        super().__init__()
        self.channels, self.filters, self.mlp_channels = list(channels), list(filters), list(mlp_channels)
        self.node_dim, self.edge_dim = node_dim + keig, edge_dim + keig
        f0 = self.filters[0]
        self.HL_init_conv = NEConvBlock(self.node_dim, self.edge_dim, f0, K, dropout_ratio)
        fin = f0
        for i, fout in enumerate(self.filters):
            for j in range(self.channels[i]):
                setattr(self, f"NEInt{i}{j}", NodeEdgeInt(d=fin, dv=fout))
                setattr(self, f"NEConv{i}{j}", NEConvBlock(fout, fout, fout, K, dropout_ratio))
                fin = fin + fout
        m_in = self.filters[-1] * 2
        for i, m_out in enumerate(self.mlp_channels):
            setattr(self, f"mlp{i}", nn.Sequential(nn.Linear(m_in, m_out), nn.BatchNorm1d(m_out),
                                                   nn.ReLU(), nn.Dropout(dropout_ratio_mlp)))
            m_in = m_out
        self.out = nn.Linear(m_in, num_classes)

    def forward(self, data):
        nb = torch.repeat_interleave(torch.arange(len(data.num_node1)), data.num_node1)
        sb = torch.repeat_interleave(torch.arange(len(data.num_edge1)), data.num_edge1)
        ei_t, ew_t, ei_s, ew_s = data.edge_index_t, data.edge_weight_t, data.edge_index_s, data.edge_weight_s
        x_t, x_s = self.HL_init_conv(data.x_t, ei_t, ew_t, data.x_s, ei_s, ew_s)
        x_t0, x_s0 = x_t, x_s
        for i in range(len(self.channels)):
            par = adj2par1(data.edge_index, x_t.shape[0], x_s.shape[0], x_t.dtype)
            D = degree(data.edge_index.view(-1), dtype=x_t.dtype)        # :624 (no num_nodes / 1e-6)
            for j in range(self.channels[i]):
                x_t, x_s = getattr(self, f"NEInt{i}{j}")(x_t0, x_s0, par, D)
                x_t, x_s = getattr(self, f"NEConv{i}{j}")(x_t, ei_t, ew_t, x_s, ei_s, ew_s)
                x_t0 = torch.cat([x_t0, x_t], -1)
                x_s0 = torch.cat([x_s0, x_s], -1)
        x = torch.cat([global_mean_pool(x_s, sb, len(data.num_edge1)),
                       global_mean_pool(x_t, nb, len(data.num_node1))], -1)
        for i in range(len(self.mlp_channels)):
            x = getattr(self, f"mlp{i}")(x)
        return self.out(x)


# --------------------------------------------------------------------------------------
# the other BASELINE.json callers: TSP pyr, CIFAR10-superpixel attpool, peptides-func attpool
# --------------------------------------------------------------------------------------
def _seg_ids(counts):
    counts = torch.as_tensor(counts)
    return torch.repeat_interleave(torch.arange(len(counts)), counts)


def _build_stack(self, K, K_init, dropout_ratio):
    """HL_init_conv + NEInt{i}{j} / NEConv{i}{j} (identical in all four model classes, e.g.
    lib/Hodge_ST_Model.py:768-802); returns the width of the dense-connection buffer."""
    f0 = self.filters[0]
    self.HL_init_conv = NEConvBlock(self.node_dim, self.edge_dim, f0, K_init, dropout_ratio)
    fin = f0
    self._stage_width = []
    for i, fout in enumerate(self.filters):
        for j in range(self.channels[i]):
            setattr(self, f"NEInt{i}{j}", NodeEdgeInt(d=fin, dv=fout))
            setattr(self, f"NEConv{i}{j}", NEConvBlock(fout, fout, fout, K, dropout_ratio))
            fin = fin + fout
        self._stage_width.append(fin)
    return fin


def _pool_positions(datas):
    """lib/Hodge_ST_Model.py:1027-1036: float cluster ids of level 0 (column 0 of x_t / x_s) offset by the
    level-1 graph sizes."""
    nb, sb = _seg_ids(datas[0].num_node1), _seg_ids(datas[0].num_edge1)
    n_ahead = torch.cumsum(torch.cat([torch.zeros(1), datas[1].num_node1.float()]), 0, dtype=torch.long)[:-1]
    s_ahead = torch.cumsum(torch.cat([torch.zeros(1), datas[1].num_edge1.float()]), 0, dtype=torch.long)[:-1]
    return (datas[0].x_t[:, 0] + n_ahead[nb]).view(-1, 1), (datas[0].x_s[:, 0] + s_ahead[sb]).view(-1, 1)


def _cluster_mean(x_t0, x_s0, pos_t, pos_s):
    """lib/Hodge_ST_Model.py:1063-1067."""
    x_t0 = scatter_mean(x_t0, pos_t.to(torch.long))
    keep = ~torch.isinf(pos_s).view(-1)
    return x_t0, scatter_mean(x_s0[keep], pos_s[keep].to(torch.long))


class _HeadMlp(nn.Module):
    def _build_head(self, num_classes, dropout_ratio_mlp):
        m_in = self.filters[-1] * 2
        for i, m_out in enumerate(self.mlp_channels):
            setattr(self, f"mlp{i}", nn.Sequential(nn.Linear(m_in, m_out), nn.BatchNorm1d(m_out),
                                                   nn.ReLU(), nn.Dropout(dropout_ratio_mlp)))
            m_in = m_out
        self.out = nn.Linear(m_in, num_classes)

    def _head(self, x):
        for i in range(len(self.mlp_channels)):
            x = getattr(self, f"mlp{i}")(x)
        return self.out(x)


class _SeqConv(nn.Module):
    """gnn.Sequential([(HodgeLaguerreConv(K=1), ..), (gnn.BatchNorm), ReLU, Dropout]) of the TSP head
    (lib/Hodge_ST_Model.py:806-817): children module_0 (conv) [, module_1 (BN)]."""

    def __init__(self, fin, fout, with_bn, dropout=0.0):
        super().__init__()
        self.module_0 = HodgeLaguerreConv(fin, fout, 1)
        if with_bn:
            self.module_1 = GraphBatchNorm(fout)
        self.with_bn, self.drop = with_bn, nn.Dropout(dropout)

    def forward(self, x, ei, ew):
        x = self.module_0(x, ei, ew)
        return self.drop(torch.relu(self.module_1(x))) if self.with_bn else x


class HL_HGCNN_TSP_dense_int3_pyr(nn.Module):
    """lib/Hodge_ST_Model.py:756-852: per-edge output; readout cat[x_s, |B1^T x_t|/2] (:848-849) -> K=1
    conv (+BN+ReLU) -> K=1 conv, times edge_mask (= data.x_s[:,1:])."""

    def __init__(self, channels=(2, 2, 2), filters=(64, 128, 256), mlp_channels=(), K=2, node_dim=2, edge_dim=1,
                 num_classes=1, dropout_ratio=0.0, dropout_ratio_mlp=0.0, keig=20):
        super().__init__()
        self.channels, self.filters, self.mlp_channels = list(channels), list(filters), list(mlp_channels)
        self.node_dim, self.edge_dim = node_dim, edge_dim
        _build_stack(self, K, K, dropout_ratio)
        m_in = self.filters[-1] * 2
        if len(self.mlp_channels) == 1:
            self.mlp = _SeqConv(m_in, self.mlp_channels[0], True, dropout_ratio)
            m_in = self.mlp_channels[0]
        self.out = _SeqConv(m_in, num_classes, False)

    def forward(self, data):
        ei_t, ew_t, ei_s, ew_s = data.edge_index_t, data.edge_weight_t, data.edge_index_s, data.edge_weight_s
        x_s, edge_mask = data.x_s[:, :1], data.x_s[:, 1:]
        x_t, x_s = self.HL_init_conv(data.x_t, ei_t, ew_t, x_s, ei_s, ew_s)
        x_t0, x_s0 = x_t, x_s
        par = adj2par1(data.edge_index, x_t.shape[0], x_s.shape[0], x_t.dtype)
        D = degree(data.edge_index.view(-1), x_t.shape[0], dtype=x_t.dtype) + 1e-6
        for i in range(len(self.channels)):
            for j in range(self.channels[i]):
                x_t, x_s = getattr(self, f"NEInt{i}{j}")(x_t0, x_s0, par, D)
                x_t, x_s = getattr(self, f"NEConv{i}{j}")(x_t, ei_t, ew_t, x_s, ei_s, ew_s)
                x_t0 = torch.cat([x_t0, x_t], -1)
                x_s0 = torch.cat([x_s0, x_s], -1)
        x_t2s = torch.sparse.mm(par.transpose(0, 1), x_t).abs() / 2
        x_s = torch.cat([x_s, x_t2s], -1)
        if len(self.mlp_channels) == 1:
            x_s = self.mlp(x_s, ei_s, ew_s)
        return self.out(x_s, ei_s, ew_s) * edge_mask, _seg_ids(data.num_edge1)


class _AttPoolBase(_HeadMlp):
    def _level(self, datas, k, x_t0, x_s0):
        d = datas[k]
        par = adj2par1(d.edge_index, x_t0.shape[0], x_s0.shape[0], x_t0.dtype)
        D = degree(d.edge_index.view(-1), x_t0.shape[0], dtype=x_t0.dtype) + 1e-6
        return par, D, d.edge_index_t, d.edge_weight_t, d.edge_index_s, d.edge_weight_s

    def _readout(self, datas, x_t, x_s, last_stage):
        d = datas[min(last_stage, 1)]
        nb, sb = _seg_ids(d.num_node1), _seg_ids(d.num_edge1)
        return torch.cat([global_mean_pool(x_s, sb, len(d.num_edge1)), global_mean_pool(x_t, nb, len(d.num_node1))], -1)


class HL_HGCNN_CIFAR10SP_dense_int3_attpool(_AttPoolBase):
    """lib/Hodge_ST_Model.py:958-1091.  As written in the reference the gate only rescales the stage
    outputs x_t / x_s (:1061-1062), which the next stage overwrites: the dense-connection buffers are
    pooled UNGATED (:1064-1067) and NEAtt receives no gradient unless pool_loc is the last stage."""

    def __init__(self, channels=(2, 2, 2), filters=(64, 128, 256), mlp_channels=(), K=2, node_dim=5, l=0.5, edge_dim=4,
                 num_classes=10, dropout_ratio=0.0, dropout_ratio_mlp=0.0, pool_loc=0, keig=10):
        super().__init__()
        self.channels, self.filters, self.mlp_channels = list(channels), list(filters), list(mlp_channels)
        self.node_dim, self.edge_dim, self.pool_loc = node_dim + keig, edge_dim + keig, pool_loc
        _build_stack(self, K, 1, dropout_ratio)
        f = self.filters[pool_loc]
        setattr(self, f"NEAtt{pool_loc}", NodeEdgeInt(d=f, dv=f, only_att=True, sigma=nn.ReLU(), l=l))
        self._build_head(num_classes, dropout_ratio_mlp)

    def forward(self, datas, if_att=False):
        pos_t, pos_s = _pool_positions(datas)
        par, D, ei_t, ew_t, ei_s, ew_s = self._level(datas, 0, datas[0].x_t, datas[0].x_s)
        x_t, x_s = self.HL_init_conv(datas[0].x_t[:, 1:], ei_t, ew_t, datas[0].x_s[:, 1:], ei_s, ew_s)
        x_t0, x_s0 = x_t, x_s
        att_t = att_s = None
        for i in range(len(self.channels)):
            for j in range(self.channels[i]):
                x_t, x_s = getattr(self, f"NEInt{i}{j}")(x_t0, x_s0, par, D)
                x_t, x_s = getattr(self, f"NEConv{i}{j}")(x_t, ei_t, ew_t, x_s, ei_s, ew_s)
                x_t0 = torch.cat([x_t0, x_t], -1)
                x_s0 = torch.cat([x_s0, x_s], -1)
            if i == self.pool_loc:
                att_t, att_s = getattr(self, f"NEAtt{i}")(x_t, x_s, par, D)
                att_t, att_s = att_t / att_t.max(), att_s / att_s.max()
                x_t, x_s = x_t * att_t, x_s * att_s
                x_t0, x_s0 = _cluster_mean(x_t0, x_s0, pos_t, pos_s)
                par, D, ei_t, ew_t, ei_s, ew_s = self._level(datas, 1, x_t0, x_s0)
        out = self._head(self._readout(datas, x_t, x_s, len(self.channels) - 1))
        return (out, att_t, att_s) if if_att else out


class HL_HGCNN_pepfunc_dense_int3_attpool(_AttPoolBase):
    """main_pepfunc_HL_HGCNN_dense_int3_attpool.py:36-168: a sigmoid gate after EVERY stage on the full
    dense-connection buffer (:131-134), cluster pooling after stage pool_loc (:137-148)."""

    def __init__(self, channels=(2, 2, 2, 2), filters=(64, 128, 256, 512), mlp_channels=(), K=2, node_dim=9, edge_dim=3,
                 num_classes=10, dropout_ratio=0.0, dropout_ratio_mlp=0.0, pool_loc=0, keig=20):
        super().__init__()
        self.channels, self.filters, self.mlp_channels = list(channels), list(filters), list(mlp_channels)
        self.node_dim, self.edge_dim, self.pool_loc = node_dim + keig, edge_dim + keig, pool_loc
        _build_stack(self, K, 1, dropout_ratio)
        for i, f in enumerate(self.filters):
            setattr(self, f"NEAtt{i}", NodeEdgeInt(d=self._stage_width[i], dv=f, only_att=True, l=0.5))
        self._build_head(num_classes, dropout_ratio_mlp)

    def forward(self, datas, if_att=False):
        pos_t, pos_s = _pool_positions(datas)
        par, D, ei_t, ew_t, ei_s, ew_s = self._level(datas, 0, datas[0].x_t, datas[0].x_s)
        x_t, x_s = self.HL_init_conv(datas[0].x_t[:, 1:], ei_t, ew_t, datas[0].x_s[:, 1:], ei_s, ew_s)
        x_t0, x_s0 = x_t, x_s
        for i in range(len(self.channels)):
            for j in range(self.channels[i]):
                x_t, x_s = getattr(self, f"NEInt{i}{j}")(x_t0, x_s0, par, D)
                x_t, x_s = getattr(self, f"NEConv{i}{j}")(x_t, ei_t, ew_t, x_s, ei_s, ew_s)
                x_t0 = torch.cat([x_t0, x_t], -1)
                x_s0 = torch.cat([x_s0, x_s], -1)
            att_t, att_s = getattr(self, f"NEAtt{i}")(x_t0, x_s0, par, D)
            x_t0, x_s0 = x_t0 * att_t, x_s0 * att_s
            if i == self.pool_loc:
                x_t0, x_s0 = _cluster_mean(x_t0, x_s0, pos_t, pos_s)
                par, D, ei_t, ew_t, ei_s, ew_s = self._level(datas, 1, x_t0, x_s0)
        out = self._head(self._readout(datas, x_t, x_s, len(self.channels) - 1))
        return (out, att_t, att_s) if if_att else out


# --------------------------------------------------------------------------------------
# eigenvector positional encodings (lib/Hodge_Dataset.py:97-112)
# --------------------------------------------------------------------------------------
def eig_pe(L, k=9):
    """`eig_pe` of the reference: scipy.linalg.eigh of the dense matrix, columns reordered by ascending eigenvalue,
    eigenvectors 1 .. k-1 (the first one -- the constant vector of a connected graph's Laplacian -- is dropped)."""
    import numpy as np
    from scipy.linalg import eigh
    vals, vecs = eigh(L.numpy() if torch.is_tensor(L) else L)
    vecs = np.real(vecs[:, vals.argsort()])
    return torch.from_numpy(vecs[:, 1:k])


def dense_operator(edge_index, edge_weight, n, dtype=torch.float32):
    """The dense matrix `dense_to_sparse` was applied to (row-major COO -> dense)."""
    return torch.zeros(n, n, dtype=dtype).index_put_((edge_index[0], edge_index[1]), edge_weight.to(dtype), accumulate=True)
