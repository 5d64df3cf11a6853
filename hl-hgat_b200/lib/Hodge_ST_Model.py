"""Callers of the hot path, kept API- and state_dict-compatible with the reference's
lib/Hodge_ST_Model.py so existing checkpoints load with strict=True.  Only the classes the
BASELINE.json configs drive are mirrored here."""
import torch
import torch.nn as nn

from .. import functional as F_hl
from ..simplex import incidence_for, operator_for
from .Hodge_Cheb_Conv import NEConv, NodeEdgeInt, adj2par1, degree, _bn_relu


class _MlpBlock(nn.Sequential):
    """nn.Sequential(Linear, BatchNorm1d, ReLU, Dropout) of lib/Hodge_ST_Model.py:596-601 with the
    BN+ReLU pair routed through the fused kernel."""

    def forward(self, x, nvalid=None):
        lin, bn, _, drop = self
        return drop(_bn_relu(bn, lin(x), 0.0, nvalid))


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
        # operators are bucketed once per batch; every layer below reuses the CSR tables
        op_t = operator_for(data.edge_index_t, data.edge_weight_t, n)
        op_s = operator_for(data.edge_index_s, data.edge_weight_s, e)
        seg_t = F_hl.Segments.from_counts(torch.as_tensor(data.num_node1, device=x_t.device), total=n)
        seg_s = F_hl.Segments.from_counts(torch.as_tensor(data.num_edge1, device=x_t.device), total=e)
        x_t, x_s = self.HL_init_conv(x_t, op_t, None, x_s, op_s, None, nv)
        x_s0, x_t0 = x_s, x_t
        inc = incidence_for(data.edge_index, n)
        D = getattr(data, "D", None)
        if D is None:
            D = inc.degree()                # = degree(edge_index.view(-1)) of :624 (no 1e-6 in this model)
        for i, _ in enumerate(self.channels):
            for j in range(self.channels[i]):
                x_t, x_s = getattr(self, f"NEInt{i}{j}")(x_t0, x_s0, inc, D, nv)
                x_t, x_s = getattr(self, f"NEConv{i}{j}")(x_t, op_t, None, x_s, op_s, None, nv)
                x_t0 = torch.cat([x_t0, x_t], dim=-1)
                x_s0 = torch.cat([x_s0, x_s], dim=-1)
        x = torch.cat((F_hl.segment_mean(x_s, seg_s), F_hl.segment_mean(x_t, seg_t)), -1)
        for i, _ in enumerate(self.mlp_channels):
            x = getattr(self, "mlp%d" % i)(x, nv_g)
        if if_final_layer:
            return x, self.out(x)
        return self.out(x)
