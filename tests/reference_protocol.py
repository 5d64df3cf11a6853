"""TEST INFRASTRUCTURE: the reference model classes' CALL PROTOCOL, with the operator layer injected.

`/root/reference` does not exist on the GPU box, so the unmodified `lib/Hodge_ST_Model.py` cannot be imported by a
`-m gpu` test.  What a drop-in has to survive is the way those classes drive the operator layer:
  * blocks built as `gnn.Sequential('x_t, edge_index_t, edge_weight_t, x_s, edge_index_s, edge_weight_s', [(module,
    'x_t, edge_index_t, edge_weight_t -> x_t'), ...])` -- children named `module_{position}`, convs called as
    `conv(x, edge_index, edge_weight)` with the int64 COO of the batch (lib/Hodge_ST_Model.py:556-590, :768-817);
  * `par_1 = adj2par1(data.edge_index, N, E)` (a torch sparse COO tensor) and `D = degree(...)` rebuilt per stage and
    handed to `NodeEdgeInt(x_t0, x_s0, par_1, D)` (:623-630, :836-844);
  * `torch.cat` dense connections, `global_mean_pool` readout (:631-636) / `torch.sparse.mm(par_1.transpose(0,1), x_t)`
    readout of the TSP model (:848-852), Python `n_batch` construction (:611-615).
These classes restate exactly that protocol (statement order and signature strings of the cited lines) and take the
operator layer (`ops`: HodgeLaguerreConv, NodeEdgeInt, adj2par1) and the PyG glue (`gnn`: Sequential, BatchNorm,
global_mean_pool; `degree`) as arguments.  tests/test_oracle_golden.py pins them, with the ORACLE operator layer on the
CPU, to the golden vectors produced by the unmodified reference (tests/golden/zinc_model.pt, models.pt);
tests/test_gpu_dropin.py then runs the same classes with the hlhgat_b200 operator layer on CUDA against the same vectors.
(In the build container tests/test_dropin_reference_sources.py additionally patches the real reference modules.)
"""
import torch
import torch.nn as nn


def _neconv(ops, gnn, fin_t, fin_s, fout, K, p):
    layers = [(ops.HodgeLaguerreConv(fin_t, fout, K=K), 'x_t, edge_index_t, edge_weight_t -> x_t'),
              (gnn.BatchNorm(fout), 'x_t -> x_t'),
              (nn.ReLU(), 'x_t -> x_t'),
              (nn.Dropout(p=p), 'x_t -> x_t'),
              (ops.HodgeLaguerreConv(fin_s, fout, K=K), 'x_s, edge_index_s, edge_weight_s -> x_s'),
              (gnn.BatchNorm(fout), 'x_s -> x_s'),
              (nn.ReLU(), 'x_s -> x_s'),
              (nn.Dropout(p=p), 'x_s -> x_s'),
              (lambda x1, x2: [x1, x2], 'x_t, x_s -> x')]
    return gnn.Sequential('x_t, edge_index_t, edge_weight_t, x_s, edge_index_s, edge_weight_s', layers)


def _stack(self, ops, gnn, K, p):
    self.HL_init_conv = _neconv(ops, gnn, self.node_dim, self.edge_dim, self.filters[0], K, p)
    fin = self.filters[0]
    for i, fout in enumerate(self.filters):
        for j in range(self.channels[i]):
            setattr(self, 'NEInt{}{}'.format(i, j), ops.NodeEdgeInt(d=fin, dv=fout))
            setattr(self, 'NEConv{}{}'.format(i, j), _neconv(ops, gnn, fout, fout, fout, K, p))
            fin = fout + fin


def _batch_vector(counts, device):
    return torch.cat([torch.tensor([i] * int(n)) for i, n in enumerate(counts)], dim=-1).to(device)


class ZincPyrProtocol(nn.Module):
    """lib/Hodge_ST_Model.py:544-646."""

    def __init__(self, ops, gnn, degree, channels, filters, mlp_channels=(), K=2, node_dim=21, edge_dim=3, num_classes=1,
                 dropout_ratio=0.0, dropout_ratio_mlp=0.0, keig=7):
        super().__init__()
        self.ops, self.gnn, self.degree = ops, gnn, degree
        self.channels, self.filters, self.mlp_channels = list(channels), list(filters), list(mlp_channels)
        self.node_dim, self.edge_dim = node_dim + keig, edge_dim + keig
        _stack(self, ops, gnn, K, dropout_ratio)
        m_in = self.filters[-1] * 2
        for i, m_out in enumerate(self.mlp_channels):
            setattr(self, 'mlp%d' % i, nn.Sequential(nn.Linear(m_in, m_out), nn.BatchNorm1d(m_out), nn.ReLU(),
                                                     nn.Dropout(dropout_ratio_mlp)))
            m_in = m_out
        self.out = nn.Linear(m_in, num_classes)

    def forward(self, data, device='cuda:0'):
        n_batch = _batch_vector(data.num_node1, device)
        s_batch = _batch_vector(data.num_edge1, device)
        x_s, edge_index_s, edge_weight_s = data.x_s, data.edge_index_s, data.edge_weight_s
        x_t, edge_index_t, edge_weight_t = data.x_t, data.edge_index_t, data.edge_weight_t
        x_t, x_s = self.HL_init_conv(x_t, edge_index_t, edge_weight_t, x_s, edge_index_s, edge_weight_s)
        x_s0, x_t0 = x_s, x_t
        for i, _ in enumerate(self.channels):
            par_1 = self.ops.adj2par1(data.edge_index, x_t.shape[0], x_s.shape[0])
            D = self.degree(data.edge_index.view(-1))
            for j in range(self.channels[i]):
                x_t, x_s = getattr(self, 'NEInt{}{}'.format(i, j))(x_t0, x_s0, par_1, D)
                x_t, x_s = getattr(self, 'NEConv{}{}'.format(i, j))(x_t, edge_index_t, edge_weight_t, x_s, edge_index_s,
                                                                    edge_weight_s)
                x_t0 = torch.cat([x_t0, x_t], dim=-1)
                x_s0 = torch.cat([x_s0, x_s], dim=-1)
        x = torch.cat((self.gnn.global_mean_pool(x_s, s_batch), self.gnn.global_mean_pool(x_t, n_batch)), -1)
        for i, _ in enumerate(self.mlp_channels):
            x = getattr(self, 'mlp%d' % i)(x)
        return self.out(x)


class TspPyrProtocol(nn.Module):
    """lib/Hodge_ST_Model.py:756-852."""

    def __init__(self, ops, gnn, degree, channels, filters, mlp_channels=(), K=2, node_dim=2, edge_dim=1, num_classes=1,
                 dropout_ratio=0.0, dropout_ratio_mlp=0.0, keig=20):
        super().__init__()
        self.ops, self.gnn, self.degree = ops, gnn, degree
        self.channels, self.filters, self.mlp_channels = list(channels), list(filters), list(mlp_channels)
        self.node_dim, self.edge_dim = node_dim, edge_dim
        _stack(self, ops, gnn, K, dropout_ratio)
        m_in = self.filters[-1] * 2
        if len(self.mlp_channels) == 1:
            layers = [(ops.HodgeLaguerreConv(m_in, self.mlp_channels[0], K=1), 'x_t, edge_index_t, edge_weight_t -> x_t'),
                      (gnn.BatchNorm(self.mlp_channels[0]), 'x_t -> x_t'),
                      (nn.ReLU(), 'x_t -> x_t'),
                      (nn.Dropout(p=dropout_ratio), 'x_t -> x_t')]
            self.mlp = gnn.Sequential('x_t, edge_index_t, edge_weight_t', layers)
            m_in = self.mlp_channels[0]
        self.out = gnn.Sequential('x_t, edge_index_t, edge_weight_t',
                                  [(ops.HodgeLaguerreConv(m_in, num_classes, K=1), 'x_t, edge_index_t, edge_weight_t -> x_t')])

    def forward(self, data, device='cuda:0'):
        s_batch = _batch_vector(data.num_edge1, device)
        x_s, edge_index_s, edge_weight_s = data.x_s[:, :1], data.edge_index_s, data.edge_weight_s
        edge_mask = data.x_s[:, 1:]
        x_t, edge_index_t, edge_weight_t = data.x_t, data.edge_index_t, data.edge_weight_t
        x_t, x_s = self.HL_init_conv(x_t, edge_index_t, edge_weight_t, x_s, edge_index_s, edge_weight_s)
        x_s0, x_t0 = x_s, x_t
        par_1 = self.ops.adj2par1(data.edge_index, x_t.shape[0], x_s.shape[0])
        D = self.degree(data.edge_index.view(-1), num_nodes=x_t.shape[0]) + 1e-6
        for i, _ in enumerate(self.channels):
            for j in range(self.channels[i]):
                x_t, x_s = getattr(self, 'NEInt{}{}'.format(i, j))(x_t0, x_s0, par_1, D)
                x_t, x_s = getattr(self, 'NEConv{}{}'.format(i, j))(x_t, edge_index_t, edge_weight_t, x_s, edge_index_s,
                                                                    edge_weight_s)
                x_t0 = torch.cat([x_t0, x_t], dim=-1)
                x_s0 = torch.cat([x_s0, x_s], dim=-1)
        x_t2s = torch.sparse.mm(par_1.transpose(0, 1), x_t).abs() / 2
        x_s = torch.cat([x_s, x_t2s], dim=-1)
        if len(self.mlp_channels) == 1:
            x_s = self.mlp(x_s, edge_index_s, edge_weight_s)
        return self.out(x_s, edge_index_s, edge_weight_s) * edge_mask, s_batch
