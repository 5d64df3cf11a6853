"""Callers of the hot path, kept API- and state_dict-compatible with the reference's
lib/Hodge_ST_Model.py so existing checkpoints load with strict=True.  Only the classes the
BASELINE.json configs drive are mirrored here."""
import contextlib
import os

import torch
import torch.nn as nn

from .. import functional as F_hl
from .. import lanes as _lanes
from ..dense_stack import new_stack
from ..simplex import Hodge1Factor, incidence_for, operator_for
from .Hodge_Cheb_Conv import NEConv, NodeEdgeInt, adj2par1, degree, _bn_relu, _epilogue_stats, node_edge_int_on_stack


_PARALLEL_BUCKETING = os.environ.get("HL_PARALLEL_BUCKETING", "1") != "0"


def _share_tables(ln, op_s, inc):
    """Tables built on the node lane that the edge lane reads too (lanes.py: record_stream at the crossing)."""
    if ln is None:
        return
    tabs = [inc.tail, inc.head, inc.rowptr, inc.edge]
    for csr in (op_s._fwd, op_s._bwd):
        if csr is not None:
            tabs += [t for t in csr if t is not None]
    if op_s.factored is not None:
        tabs.append(op_s.factored.edge_scale)
        fi = op_s.factored.incidence
        tabs += [fi.tail, fi.head, fi.rowptr, fi.edge]
    ln.to_edge(*tabs)


class _MlpBlock(nn.Sequential):
    """nn.Sequential(Linear, BatchNorm1d, ReLU, Dropout) of lib/Hodge_ST_Model.py:596-601 with the
    BN+ReLU pair routed through the fused kernel."""

    def forward(self, x, nvalid=None):
        lin, bn, _, drop = self
        with _epilogue_stats(bn, nvalid) as tiles:
            h = F_hl.linear(x, lin.weight, lin.bias)
        return drop(_bn_relu(bn, h, 0.0, nvalid, tiles=tiles))


class HL_HGCNN_zinc_dense_int3_pyr(nn.Module):
    """Reference lib/Hodge_ST_Model.py:544-646."""

    def __init__(self, channels=[2, 2, 2, 2], filters=[64, 128, 256, 512], mlp_channels=[], K=2, node_dim=21,
                 edge_dim=3, num_classes=1, dropout_ratio=0.0, dropout_ratio_mlp=0.0, keig=7):
        super().__init__()
        self.channels = channels
        self.filters = filters
        self.mlp_channels = mlp_channels
        self.node_dim = node_dim + keig
        self.edge_dim = edge_dim + keig
        self.initial_channel = self.filters[0]
        self.HL_init_conv = NEConv(self.node_dim, self.edge_dim, self.initial_channel, K, dropout_ratio)
        gcn_insize = self.initial_channel
        for i, gcn_outsize in enumerate(self.filters):
            for j in range(self.channels[i]):
                setattr(self, f"NEInt{i}{j}", NodeEdgeInt(d=gcn_insize, dv=gcn_outsize))
                setattr(self, f"NEConv{i}{j}", NEConv(gcn_outsize, gcn_outsize, gcn_outsize, K, dropout_ratio))
                gcn_insize = gcn_outsize + gcn_insize
        mlp_insize = self.filters[-1] * 2
        for i, mlp_outsize in enumerate(mlp_channels):
            setattr(self, "mlp%d" % i, _MlpBlock(nn.Linear(mlp_insize, mlp_outsize), nn.BatchNorm1d(mlp_outsize),
                                                 nn.ReLU(), nn.Dropout(dropout_ratio_mlp)))
            mlp_insize = mlp_outsize
        self.out = nn.Linear(mlp_insize, num_classes)

    def forward(self, data, device="cuda:0", if_final_layer=False):
        x_s, x_t = data.x_s, data.x_t
        n, e = x_t.shape[0], x_s.shape[0]
        # fixed-capacity (CUDA-graph) batches carry device-side valid counts; rows beyond are padding
        nv = (getattr(data, "n_valid_nodes", None), getattr(data, "n_valid_edges", None))
        nv_g = getattr(data, "n_valid_graphs", None)
        with _lanes.open_lanes(x_t.device) as ln:      # ln is None unless lanes.enable_lanes(): single stream
            # operators are bucketed once per batch; every layer below reuses the CSR tables.  With lanes the edge lane
            # buckets its own operator (only it reads L1) while the node lane does L0, the incidence tables and the
            # segments: the serial prefix of the step is the longer of the two, not their sum.  The edge lane first
            # meets the incidence tables after the exchange of the first NodeEdgeInt, which orders it behind them.
            with (ln.edge_ctx() if ln is not None and _PARALLEL_BUCKETING else contextlib.nullcontext()):
                op_s = operator_for(data.edge_index_s, data.edge_weight_s, e)
            op_t = operator_for(data.edge_index_t, data.edge_weight_t, n)
            inc = incidence_for(data.edge_index, n)
            D = getattr(data, "D", None)
            if D is None:
                D = inc.degree()            # = degree(edge_index.view(-1)) of :624 (no 1e-6 in this model)
            seg_t = F_hl.Segments.from_counts(torch.as_tensor(data.num_node1, device=x_t.device), total=n, ghost_last=nv_g is not None)
            seg_s = F_hl.Segments.from_counts(torch.as_tensor(data.num_edge1, device=x_t.device), total=e, ghost_last=nv_g is not None)
            if ln is not None and _PARALLEL_BUCKETING:
                ln._mark(ln.edge, [inc.tail, inc.head, inc.rowptr, inc.edge, D])     # read on the edge lane later (no wait here)
            else:
                _share_tables(ln, op_s, inc)
            last = (len(self.channels) - 1, self.channels[-1] - 1)
            # dense connections in preallocated buffers: no torch.cat, every block transferred once (dense_stack.py);
            # the block of the last layer is never read (:627-636) and stays out of the buffers
            width = self.initial_channel + sum(f * c for f, c in zip(self.filters, self.channels)) - self.filters[-1]
            stack = new_stack(n, e, width, inc, D, x_t.device)
            x_t, x_s = self.HL_init_conv(x_t, op_t, None, x_s, op_s, None, nv, stack=stack)
            for i, _ in enumerate(self.channels):
                for j in range(self.channels[i]):
                    x_t, x_s = node_edge_int_on_stack(getattr(self, f"NEInt{i}{j}"), stack, nv)
                    x_t, x_s = getattr(self, f"NEConv{i}{j}")(x_t, op_t, None, x_s, op_s, None, nv,
                                                              stack=None if (i, j) == last else stack)
            stack.close()
            if ln is not None:
                ln.to_node(x_s)
        x = torch.cat((F_hl.segment_mean(x_s, seg_s), F_hl.segment_mean(x_t, seg_t)), -1)
        for i, _ in enumerate(self.mlp_channels):
            x = getattr(self, "mlp%d" % i)(x, nv_g)
        if if_final_layer:
            return x, self.out(x)
        return self.out(x)


# ---------------------------------------------------------------------------------------------
# the other BASELINE.json callers (TSP pyr, CIFAR10-superpixel attpool, peptides-func attpool)
# ---------------------------------------------------------------------------------------------
def _build_stack(self, K, K_init, dropout_ratio):
    """HL_init_conv + NEInt{i}{j} / NEConv{i}{j}: identical in all model classes of the reference
    (e.g. lib/Hodge_ST_Model.py:768-802).  Records the dense-connection width after every stage."""
    f0 = self.filters[0]
    self.HL_init_conv = NEConv(self.node_dim, self.edge_dim, f0, K_init, dropout_ratio)
    fin = f0
    self._stage_width = []
    for i, fout in enumerate(self.filters):
        for j in range(self.channels[i]):
            setattr(self, f"NEInt{i}{j}", NodeEdgeInt(d=fin, dv=fout))
            setattr(self, f"NEConv{i}{j}", NEConv(fout, fout, fout, K, dropout_ratio))
            fin = fin + fout
        self._stage_width.append(fin)


class _Level:
    """Device-side tables of one coarsening level of a batch: operators, incidence, degree, segments."""

    def __init__(self, d, n, e, eps, device):
        self.n, self.e = n, e
        # batches built on the GPU (lib.Hodge_Dataset.two_level_batch_from_graphs) carry their CSR / incidence tables
        self.op_t = getattr(d, "op_t", None) or operator_for(d.edge_index_t.to(device), d.edge_weight_t.to(device), n)
        self.op_s = getattr(d, "op_s", None) or operator_for(d.edge_index_s.to(device), d.edge_weight_s.to(device), e)
        self.inc = getattr(d, "incidence", None) or incidence_for(d.edge_index.to(device), n)
        if F_hl.factored_hodge1_enabled() and self.op_s.factored is None:
            # opt-in: the caller asserts edge_index_s / edge_weight_s = 2 B1^T B1 / lambda_max (lib/Hodge_Dataset.py:456)
            self.op_s.factored = Hodge1Factor.from_operator(self.op_s, self.inc)
        D = getattr(d, "D", None)
        self.D = (self.inc.degree() + eps) if D is None else D.to(device)   # degree(edge_index.view(-1), n) + 1e-6
        self.nv = (getattr(d, "n_valid_nodes", None), getattr(d, "n_valid_edges", None))
        self.nv_g = getattr(d, "n_valid_graphs", None)
        self._d, self._device = d, device

    def segments(self, readout=False):
        """Per-graph row segments; readout=True: for the mean readout of a padded batch the ghost graph's segment is
        left empty (Segments.from_counts)."""
        d, dev = self._d, self._device
        ghost = readout and self.nv_g is not None
        return (F_hl.Segments.from_counts(torch.as_tensor(d.num_node1, device=dev), total=self.n, ghost_last=ghost),
                F_hl.Segments.from_counts(torch.as_tensor(d.num_edge1, device=dev), total=self.e, ghost_last=ghost))


def _stage(self, i, lv, stack, keep_last=True):
    """One stage of NEInt / NEConv layers with dense connections held in `stack` (dense_stack.py).  With lanes enabled
    (lanes.py) the edge chain is issued on the edge lane; the stage hands everything back to the node lane at its end
    (gates, pooling and the head run there).  keep_last=False: the block of the stage's last layer is not appended
    (nothing reads the buffers afterwards)."""
    x_t = x_s = None
    ln = _lanes.active()
    if ln is not None:
        _share_tables(ln, lv.op_s, lv.inc)
        ln.edge.wait_stream(ln.node)
    nl = self.channels[i]
    for j in range(nl):
        x_t, x_s = node_edge_int_on_stack(getattr(self, f"NEInt{i}{j}"), stack, lv.nv)
        x_t, x_s = getattr(self, f"NEConv{i}{j}")(x_t, lv.op_t, None, x_s, lv.op_s, None, lv.nv,
                                                  stack=stack if (keep_last or j + 1 < nl) else None)
    if ln is not None:
        ln.to_node(x_s)
    return x_t, x_s


def _stack_width(self, first, stages, drop_last):
    """Columns a stack needs: `first` (its initial block) + the layers of `stages`, minus the very last block if
    nothing reads the buffers after it."""
    w = first + sum(self.channels[i] * self.filters[i] for i in stages)
    return w - (self.filters[stages[-1]] if drop_last and stages else 0)


class _SeqConv(nn.Module):
    """gnn.Sequential([HodgeLaguerreConv(K=1) [, gnn.BatchNorm, ReLU, Dropout]]) of the TSP head
    (lib/Hodge_ST_Model.py:806-817): children module_0 (conv) [, module_1 (BN)]."""

    def __init__(self, fin, fout, with_bn, dropout=0.0):
        super().__init__()
        from .Hodge_Cheb_Conv import HodgeLaguerreConv, GraphBatchNorm
        self.module_0 = HodgeLaguerreConv(fin, fout, K=1)
        if with_bn:
            self.module_1 = GraphBatchNorm(fout)
        self.with_bn, self.p = with_bn, dropout

    def forward(self, x, op, nvalid=None):
        x = self.module_0(x, op, None)
        if self.with_bn:
            x = self.module_1.forward_act(x, 0.0, nvalid)
            if self.p > 0.0:
                x = torch.nn.functional.dropout(x, self.p, self.training)
        return x


class HL_HGCNN_TSP_dense_int3_pyr(nn.Module):
    """Reference lib/Hodge_ST_Model.py:756-852 (per-edge output, readout through |B1^T x_t|/2 :848)."""

    def __init__(self, channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[], K=2, node_dim=2, edge_dim=1,
                 num_classes=1, dropout_ratio=0.0, dropout_ratio_mlp=0.0, keig=20):
        super().__init__()
        self.channels, self.filters, self.mlp_channels = channels, filters, mlp_channels
        self.node_dim, self.edge_dim = node_dim, edge_dim
        self.initial_channel = self.filters[0]
        _build_stack(self, K, K, dropout_ratio)
        mlp_insize = self.filters[-1] * 2
        if len(self.mlp_channels) == 1:
            self.mlp = _SeqConv(mlp_insize, self.mlp_channels[0], True, dropout_ratio)
            mlp_insize = self.mlp_channels[0]
        self.out = _SeqConv(mlp_insize, num_classes, False)

    def forward(self, data, device="cuda:0"):
        x_t = data.x_t
        x_s, edge_mask = data.x_s[:, :1], data.x_s[:, 1:]
        lv = _Level(data, x_t.shape[0], x_s.shape[0], 1e-6, x_t.device)
        with _lanes.open_lanes(x_t.device) as ln:
            _share_tables(ln, lv.op_s, lv.inc)
            stages = list(range(len(self.channels)))
            stack = new_stack(lv.n, lv.e, _stack_width(self, self.initial_channel, stages, True), lv.inc, lv.D, x_t.device)
            x_t, x_s = self.HL_init_conv(x_t, lv.op_t, None, x_s, lv.op_s, None, lv.nv, stack=stack)
            for i in stages:
                x_t, x_s = _stage(self, i, lv, stack, keep_last=i != stages[-1])
            stack.close()
        x_s = torch.cat([x_s, F_hl.boundary_absdiff(x_t, lv.inc)], dim=-1)
        if len(self.mlp_channels) == 1:
            x_s = self.mlp(x_s, lv.op_s, lv.nv[1])
        s_batch = lv.segments()[1].owner.long()
        return self.out(x_s, lv.op_s) * edge_mask, s_batch


class _AttPool(nn.Module):
    def _build_head(self, num_classes, dropout_ratio_mlp):
        mlp_insize = self.filters[-1] * 2
        for i, mlp_outsize in enumerate(self.mlp_channels):
            setattr(self, "mlp%d" % i, _MlpBlock(nn.Linear(mlp_insize, mlp_outsize), nn.BatchNorm1d(mlp_outsize),
                                                 nn.ReLU(), nn.Dropout(dropout_ratio_mlp)))
            mlp_insize = mlp_outsize
        self.out = nn.Linear(mlp_insize, num_classes)

    @staticmethod
    def _positions(datas, lv0, n1, e1, device):
        """lib/Hodge_ST_Model.py:1027-1036: float cluster ids (column 0 of x_t / x_s) offset by the level-1
        sizes of the preceding graphs; bucketed once into pooling segments (+inf = edge inside a cluster)."""
        seg_n, seg_e = lv0.segments()
        n_ahead = torch.cumsum(torch.as_tensor(datas[1].num_node1, device=device), 0) - torch.as_tensor(datas[1].num_node1, device=device)
        s_ahead = torch.cumsum(torch.as_tensor(datas[1].num_edge1, device=device), 0) - torch.as_tensor(datas[1].num_edge1, device=device)
        pos_t = (datas[0].x_t[:, 0] + n_ahead[seg_n.owner.long()]).view(-1, 1)
        pos_s = (datas[0].x_s[:, 0] + s_ahead[seg_e.owner.long()]).view(-1, 1)
        return F_hl.Segments.from_index(pos_t, nrows=n1), F_hl.Segments.from_index(pos_s, nrows=e1)

    def _head(self, x_t, x_s, lv):
        seg_n, seg_e = lv.segments(readout=True)
        x = torch.cat((F_hl.segment_mean(x_s, seg_e), F_hl.segment_mean(x_t, seg_n)), -1)
        for i, _ in enumerate(self.mlp_channels):
            x = getattr(self, "mlp%d" % i)(x, lv.nv_g)
        return x


class HL_HGCNN_CIFAR10SP_dense_int3_attpool(_AttPool):
    """Reference lib/Hodge_ST_Model.py:958-1091.  As in the reference, the ReLU gate (normalised by its
    batch maximum, :1059-1060) only rescales the stage outputs x_t / x_s, which the next stage overwrites;
    the dense-connection buffers are pooled ungated (:1064-1067)."""

    def __init__(self, channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[], K=2, node_dim=5, l=0.5, edge_dim=4,
                 num_classes=10, dropout_ratio=0.0, dropout_ratio_mlp=0.0, pool_loc=0, keig=10):
        super().__init__()
        self.channels, self.filters, self.mlp_channels = channels, filters, mlp_channels
        self.node_dim, self.edge_dim = node_dim + keig, edge_dim + keig
        self.initial_channel = self.filters[0]
        self.pool_loc = pool_loc
        _build_stack(self, K, 1, dropout_ratio)
        f = self.filters[pool_loc]
        setattr(self, f"NEAtt{pool_loc}", NodeEdgeInt(d=f, dv=f, only_att=True, sigma=nn.ReLU(), l=l))
        self._build_head(num_classes, dropout_ratio_mlp)

    def forward(self, datas, device="cuda:0", if_final_layer=False, if_att=False):
        d0, d1 = datas[0], datas[1]
        dev = d0.x_t.device
        lv = _Level(d0, d0.x_t.shape[0], d0.x_s.shape[0], 1e-6, dev)
        n1, e1 = d1.x_t.shape[0], d1.x_s.shape[0]
        seg_pt, seg_ps = self._positions(datas, lv, n1, e1, dev)
        with _lanes.open_lanes(dev) as ln:
            _share_tables(ln, lv.op_s, lv.inc)
            ns = len(self.channels)
            before, after = list(range(self.pool_loc + 1)), list(range(self.pool_loc + 1, ns))
            stack = new_stack(lv.n, lv.e, _stack_width(self, self.initial_channel, before, False), lv.inc, lv.D, dev)
            x_t, x_s = self.HL_init_conv(d0.x_t[:, 1:], lv.op_t, None, d0.x_s[:, 1:], lv.op_s, None, lv.nv, stack=stack)
            att_t = att_s = None
            for i in range(ns):
                x_t, x_s = _stage(self, i, lv, stack, keep_last=i == self.pool_loc or i != ns - 1)
                if i == self.pool_loc:
                    att_t, att_s = getattr(self, "NEAtt%d" % i)(x_t, x_s, lv.inc, lv.D)
                    att_t = att_t / att_t.max()
                    att_s = att_s / att_s.max()
                    x_t, x_s = x_t * att_t, x_s * att_s
                    x_t0 = F_hl.segment_mean(stack.whole("t"), seg_pt)
                    x_s0 = F_hl.segment_mean(stack.whole("s"), seg_ps)
                    lv = _Level(d1, n1, e1, 1e-6, dev)
                    stack.close()
                    stack = new_stack(n1, e1, _stack_width(self, x_t0.shape[1], after, True), lv.inc, lv.D, dev)
                    stack.publish(x_t0, x_s0)
            stack.close()
        x = self._head(x_t, x_s, lv)
        if if_final_layer:
            return x, self.out(x)
        if if_att:
            return self.out(x), att_t, att_s
        return self.out(x)


class HL_HGCNN_pepfunc_dense_int3_attpool(_AttPool):
    """Reference main_pepfunc_HL_HGCNN_dense_int3_attpool.py:36-168 (the class the peptides-func script
    trains): a sigmoid gate on the whole dense-connection buffer after EVERY stage (:131-134); at stage
    pool_loc the gate multiply is fused into the cluster-mean kernel (:137-143)."""

    def __init__(self, channels=[2, 2, 2, 2], filters=[64, 128, 256, 512], mlp_channels=[], K=2, node_dim=9, edge_dim=3,
                 num_classes=10, dropout_ratio=0.0, dropout_ratio_mlp=0.0, pool_loc=0, keig=20):
        super().__init__()
        self.channels, self.filters, self.mlp_channels = channels, filters, mlp_channels
        self.node_dim, self.edge_dim = node_dim + keig, edge_dim + keig
        self.initial_channel = self.filters[0]
        self.pool_loc = pool_loc
        self.relu = nn.ReLU()
        _build_stack(self, K, 1, dropout_ratio)
        for i, f in enumerate(self.filters):
            setattr(self, f"NEAtt{i}", NodeEdgeInt(d=self._stage_width[i], dv=f, only_att=True, l=0.5))
        self._build_head(num_classes, dropout_ratio_mlp)

    def forward(self, datas, device="cuda:0", if_att=False, if_final_layer=False):
        d0, d1 = datas[0], datas[1]
        dev = d0.x_t.device
        lv = _Level(d0, d0.x_t.shape[0], d0.x_s.shape[0], 1e-6, dev)
        n1, e1 = d1.x_t.shape[0], d1.x_s.shape[0]
        seg_pt, seg_ps = self._positions(datas, lv, n1, e1, dev)
        with _lanes.open_lanes(dev) as ln:
            _share_tables(ln, lv.op_s, lv.inc)
            last = len(self.channels) - 1
            # the sigmoid gate rescales the WHOLE dense-connection buffer after every stage (:131-134), so every stage
            # starts a fresh stack from the gated (and, at pool_loc, pooled) buffers
            stack = new_stack(lv.n, lv.e, _stack_width(self, self.initial_channel, [0], last == 0 and not if_att), lv.inc, lv.D, dev)
            x_t, x_s = self.HL_init_conv(d0.x_t[:, 1:], lv.op_t, None, d0.x_s[:, 1:], lv.op_s, None, lv.nv, stack=stack)
            for i, _ in enumerate(self.channels):
                x_t, x_s = _stage(self, i, lv, stack, keep_last=i != last or if_att)
                if i == last and not if_att:
                    stack.close()
                    break                      # the last gate only rescales buffers nothing reads (:131-134 then :150)
                x_t0, x_s0 = stack.whole("t"), stack.whole("s")
                stack.close()
                att_t, att_s = getattr(self, "NEAtt%d" % i)(x_t0, x_s0, lv.inc, lv.D)
                if i == last:
                    break
                if i == self.pool_loc:
                    x_t0 = F_hl.segment_mean(x_t0, seg_pt, att_t)
                    x_s0 = F_hl.segment_mean(x_s0, seg_ps, att_s)
                    lv = _Level(d1, n1, e1, 1e-6, dev)
                else:
                    x_t0, x_s0 = x_t0 * att_t, x_s0 * att_s
                stack = new_stack(lv.n, lv.e, _stack_width(self, x_t0.shape[1], [i + 1], i + 1 == last and not if_att),
                                  lv.inc, lv.D, dev)
                stack.publish(x_t0, x_s0)
        x = self._head(x_t, x_s, lv)
        if if_att:
            return self.out(x), att_t, att_s
        elif if_final_layer:
            return x, self.out(x)
        return self.out(x)
