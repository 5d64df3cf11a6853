"""Parity on the shapes of the other BASELINE.json configs (peptides-func, CIFAR10-superpixel, TSP):
polynomial convs of the orders those scripts use, the NodeEdgeInt value/gate paths and the
attention-pooling block on a coarsened level, against the CPU oracle (rtol 1e-4)."""
import numpy as np
import pytest
import torch

import hlhgat_b200 as H
from hlhgat_b200 import functional as F_hl
from hlhgat_b200.synthetic import make_batch, batch_to
from oracle import hodge_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close(a, b, rtol=1e-4, atol=None):
    a, b = a.detach().cpu(), b.detach().cpu()
    atol = 1e-4 * float(b.abs().max()) if atol is None else atol
    assert a.shape == b.shape
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {(a - b).abs().max().item():.3e} (scale {b.abs().max():.3e})"


@pytest.mark.parametrize("shape,batch,K,width", [("peptides", 6, 6, 64), ("cifar", 4, 4, 64), ("tsp", 2, 4, 32),
                                                 ("tsp", 1, 2, 128), ("zinc", 40, 6, 256)])
@pytest.mark.parametrize("family", ["laguerre", "cheb"])
def test_conv_fwd_bwd_on_config_shapes(shape, batch, K, width, family):
    torch.manual_seed(K * width)
    b = make_batch(shape, batch, seed=5)
    for side in ("t", "s"):
        ei, ew = getattr(b, f"edge_index_{side}"), getattr(b, f"edge_weight_{side}")
        r = b.x_t.shape[0] if side == "t" else b.x_s.shape[0]
        oc = (O.HodgeLaguerreConv if family == "laguerre" else O.HodgeChebConv)(width, width, K)
        with torch.no_grad():
            oc.bias.normal_()
        gc = (H.HodgeLaguerreConv if family == "laguerre" else H.HodgeChebConv)(width, width, K).to(DEV)
        gc.load_state_dict(oc.state_dict())
        x = torch.randn(r, width) * 0.5
        w = torch.randn(r, width) / r ** 0.5
        xo = x.clone().requires_grad_(True)
        yo = oc(xo, ei, ew)
        go = torch.autograd.grad((yo * w).sum(), [xo] + list(oc.parameters()))
        xg = x.to(DEV).requires_grad_(True)
        yg = gc(xg, ei.to(DEV), ew.to(DEV))
        gg = torch.autograd.grad((yg * w.to(DEV)).sum(), [xg] + list(gc.parameters()))
        close(yg, yo)
        for a, c in zip(gg, go):
            close(a, c)


def _coarsen(b):
    """A deterministic pairing of consecutive nodes inside each graph (stand-in for graclus + MLGC,
    lib/Hodge_Dataset.py:241-295): float cluster ids for nodes, +inf for edges inside a cluster, coarse
    edge ids in first-appearance order otherwise."""
    nb = torch.repeat_interleave(torch.arange(len(b.num_node1)), b.num_node1)
    first = torch.cat([torch.zeros(1, dtype=torch.long), b.num_node1.cumsum(0)[:-1]])
    local = torch.arange(nb.numel()) - first[nb]
    ncl = (b.num_node1 + 1) // 2
    cl_first = torch.cat([torch.zeros(1, dtype=torch.long), ncl.cumsum(0)[:-1]])
    pos_t = (cl_first[nb] + local // 2)
    a, c = pos_t[b.edge_index[0]], pos_t[b.edge_index[1]]
    key = torch.minimum(a, c) * int(pos_t.max() + 1) + torch.maximum(a, c)
    pos_s = torch.full((b.edge_index.shape[1],), float("inf"))
    seen = {}
    for i in range(key.numel()):
        if a[i] == c[i]:
            continue
        k = int(key[i])
        if k not in seen:
            seen[k] = len(seen)
        pos_s[i] = float(seen[k])
    return pos_t.float().view(-1, 1), pos_s.view(-1, 1)


@pytest.mark.parametrize("shape,batch,sigma,lam", [("peptides", 5, "sigmoid", 0.5), ("cifar", 3, "relu", 0.5)])
def test_attention_gate_and_pool_on_config_shapes(shape, batch, sigma, lam):
    torch.manual_seed(7)
    b = make_batch(shape, batch, seed=9)
    n, e, d = b.x_t.shape[0], b.x_s.shape[0], 64
    pos_t, pos_s = _coarsen(b)
    sig = torch.nn.Sigmoid() if sigma == "sigmoid" else torch.nn.ReLU()
    oa = O.NodeEdgeInt(d=d, dk=32, only_att=True, sigma=sig, l=lam)
    ga = H.NodeEdgeInt(d=d, dk=32, only_att=True, sigma=sig, l=lam).to(DEV)
    ga.load_state_dict(oa.state_dict())
    x_t, x_s = torch.randn(n, d) * 0.3, torch.randn(e, d) * 0.3
    D = O.degree(b.edge_index.view(-1), n) + 1e-6
    xo_t, xo_s = x_t.clone().requires_grad_(True), x_s.clone().requires_grad_(True)
    a_t, a_s = oa(xo_t, xo_s, O.adj2par1(b.edge_index, n, e), D)
    po_t, po_s = O.attention_pool(xo_t, xo_s, a_t, a_s, pos_t, pos_s)
    wt, ws = torch.randn_like(po_t), torch.randn_like(po_s)
    go = torch.autograd.grad((po_t * wt).sum() + (po_s * ws).sum(), [xo_t, xo_s] + list(oa.parameters()))
    xg_t, xg_s = x_t.to(DEV).requires_grad_(True), x_s.to(DEV).requires_grad_(True)
    par = H.adj2par1(b.edge_index.to(DEV), n, e)
    g_t, g_s = ga(xg_t, xg_s, par, D.to(DEV))
    close(g_t, a_t)
    close(g_s, a_s)
    pg_t = F_hl.segment_mean(xg_t, F_hl.Segments.from_index(pos_t.to(DEV)), g_t)
    pg_s = F_hl.segment_mean(xg_s, F_hl.Segments.from_index(pos_s.to(DEV)), g_s)
    close(pg_t, po_t)
    close(pg_s, po_s)
    gg = torch.autograd.grad((pg_t * wt.to(DEV)).sum() + (pg_s * ws.to(DEV)).sum(), [xg_t, xg_s] + list(ga.parameters()))
    for a, c in zip(gg, go):
        close(a, c)


def test_hl_filter_block_vs_oracle_composition():
    """HL_filter (lib/Hodge_Cheb_Conv.py:117-188) = MSI + NEConv(LeakyReLU 0.1) with dense connections."""
    torch.manual_seed(1)
    b = make_batch("peptides", 4, seed=3)
    n, e = b.x_t.shape[0], b.x_s.shape[0]
    filt = H.HL_filter(channels=2, filters=32, K=3, node_dim=16, edge_dim=16).to(DEV).train()
    x_t, x_s = torch.randn(n, 16), torch.randn(e, 16)
    D = O.degree(b.edge_index.view(-1), n) + 1e-6
    d = batch_to(b, DEV)
    par = H.adj2par1(d.edge_index, n, e)
    y_t, y_s = filt(x_t.to(DEV), d.edge_index_t, d.edge_weight_t, x_s.to(DEV), d.edge_index_s, d.edge_weight_s, par, D.to(DEV))
    assert y_t.shape == (n, 16 + 2 * 32) and y_s.shape == (e, 16 + 2 * 32)
    # oracle composition with the same parameters
    xt0, xs0 = x_t, x_s
    opar = O.adj2par1(b.edge_index, n, e)
    for j in range(2):
        msi = O.NodeEdgeInt(d=xt0.shape[1], dv=32)
        msi.load_state_dict(getattr(filt, f"MSI{j}").state_dict())
        msi.train()
        blk = O.NEConvBlock(32, 32, 32, 3, act=torch.nn.LeakyReLU(0.1))
        blk.load_state_dict(getattr(filt, f"NEConv{j}").state_dict())
        blk.train()
        a, c = msi(xt0, xs0, opar, D)
        a, c = blk(a, b.edge_index_t, b.edge_weight_t, c, b.edge_index_s, b.edge_weight_s)
        xt0, xs0 = torch.cat([xt0, a], -1), torch.cat([xs0, c], -1)
    close(y_t, xt0)
    close(y_s, xs0)
