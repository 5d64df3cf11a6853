"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors
recorded from the unmodified reference modules.  Floating point bar (BASELINE.json north_star):
rtol 1e-4 on outputs and gradients; the SpMM / transfer kernels are additionally bit-exact."""
from types import SimpleNamespace

import pytest
import torch

import hlhgat_b200 as H
from hlhgat_b200 import functional as F_hl
from hlhgat_b200 import _native as N
from hlhgat_b200.simplex import CsrOperator, Incidence, csr_from_coo
from hlhgat_b200.synthetic import make_batch, batch_to
from oracle import hodge_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-4
DEV = "cuda:0"


def close(a, b, rtol=RTOL, atol=1e-5):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {(a - b).abs().max().item():.3e}"


# ------------------------------------------------------------------------------------------
def test_csr_from_coo_matches_stable_sort():
    g = torch.Generator().manual_seed(0)
    nrows, nnz = 37, 500
    row = torch.randint(0, nrows, (nnz,), generator=g)
    col = torch.randint(0, 50, (nnz,), generator=g)
    val = torch.randn(nnz, generator=g)
    rp, ci, vv, perm = csr_from_coo(row.to(DEV), col.to(DEV), val.to(DEV), nrows, want_perm=True)
    order = torch.sort(row, stable=True)[1]
    assert torch.equal(perm.cpu().long(), order)
    assert torch.equal(ci.cpu().long(), col[order]) and torch.equal(vv.cpu(), val[order])
    counts = torch.bincount(row, minlength=nrows)
    assert torch.equal(rp.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))
    # tie-break by column, rows out of range dropped, float ids with +inf dropped
    rp2, ci2, _, _ = csr_from_coo(row.to(DEV), col.to(DEV), None, nrows, tie=N.HL_TIE_COLUMN)
    key = row * 1000 + col
    assert torch.equal(ci2.cpu().long(), col[torch.sort(key, stable=True)[1]])
    rf = row.float()
    rf[::7] = float("inf")
    rp3, ci3, _, _ = csr_from_coo(rf.to(DEV), None, None, nrows, row_is_float=True)
    keep = torch.isfinite(rf)
    kept_order = torch.arange(nnz)[keep][torch.sort(row[keep], stable=True)[1]]
    assert int(rp3[-1]) == int(keep.sum())
    assert torch.equal(ci3.cpu().long()[: int(rp3[-1])], kept_order)
    # empty input
    rp4, _, _, _ = csr_from_coo(torch.zeros(0, dtype=torch.long, device=DEV), None, None, 5)
    assert rp4.cpu().tolist() == [0] * 6


@pytest.mark.parametrize("idx", range(21))
def test_conv_vs_golden_reference(idx):
    c = load_golden("conv.pt")["cases"][idx]
    cls = H.HodgeLaguerreConv if c["family"] == "laguerre" else H.HodgeChebConv
    conv = cls(c["fin"], c["fout"], c["K"]).to(DEV)
    conv.load_state_dict(c["state"], strict=True)
    x = c["x"].to(DEV).requires_grad_(True)
    y = conv(x, c["edge_index"].to(DEV), c["edge_weight"].to(DEV))
    close(y, c["y"])
    g = torch.autograd.grad((y * c["wsum"].to(DEV)).sum(), [x] + list(conv.parameters()))
    close(g[0], c["gx"])
    for (n, _), t in zip(conv.named_parameters(), g[1:]):
        close(t, c["gp"][n])


@pytest.mark.parametrize("family", ["laguerre", "cheb"])
@pytest.mark.parametrize("width", [64, 128, 256, 28, 10, 7, 300])
def test_poly_basis_bit_exact_vs_oracle(family, width):
    """T_k computed by the SpMM kernels equals the CPU reference recurrence BIT FOR BIT."""
    b = make_batch("zinc", 24, seed=3)
    K = 5
    for ei, ew, r in ((b.edge_index_t, b.edge_weight_t, b.x_t.shape[0]), (b.edge_index_s, b.edge_weight_s, b.x_s.shape[0])):
        x = torch.randn(r, width)
        op = CsrOperator(ei.to(DEV), ew.to(DEV), r)
        (t,) = F_hl.poly_basis_fwd(F_hl._FAMILY[family], K, [op], [x.to(DEV)], width)
        t0, t1 = x, None
        for k in range(K - 1):
            if family == "laguerre":
                if k == 0:
                    t1 = x - O.propagate(x, ei, ew)
                else:
                    t2 = (-O.propagate(t1, ei, ew) + (2 * k + 1) * t1 - k * t0) / (k + 1)
                    t0, t1 = t1, t2
            else:
                if k == 0:
                    t1 = O.propagate(x, ei, ew)
                else:
                    t2 = 2. * O.propagate(t1, ei, ew) - t0
                    t0, t1 = t1, t2
            assert torch.equal(t[k].cpu(), t1), f"order {k + 1}: max err {(t[k].cpu() - t1).abs().max()}"


def test_fused_two_operator_launch_equals_separate():
    b = make_batch("zinc", 16, seed=5)
    n, e = b.x_t.shape[0], b.x_s.shape[0]
    op_t = CsrOperator(b.edge_index_t.to(DEV), b.edge_weight_t.to(DEV), n)
    op_s = CsrOperator(b.edge_index_s.to(DEV), b.edge_weight_s.to(DEV), e)
    xt, xs = torch.randn(n, 64, device=DEV), torch.randn(e, 64, device=DEV)
    both = F_hl.poly_basis_fwd(N.HL_LAGUERRE, 4, [op_t, op_s], [xt, xs], 64)
    (a,) = F_hl.poly_basis_fwd(N.HL_LAGUERRE, 4, [op_t], [xt], 64)
    (c,) = F_hl.poly_basis_fwd(N.HL_LAGUERRE, 4, [op_s], [xs], 64)
    assert torch.equal(both[0], a) and torch.equal(both[1], c)


def test_conv_3d_input_and_fastconv():
    for c in load_golden("fastconv.pt"):
        conv = H.HodgeLaguerreFastConv(5, 6, c["K"]).to(DEV)
        conv.load_state_dict(c["state"])
        close(conv(c["x"].to(DEV), (c["edge_index"].to(DEV), c["edge_weight"].to(DEV))), c["y"])
        # gradients of the quirk path against the oracle's autograd
        oc = O.HodgeLaguerreFastConv(5, 6, c["K"])
        oc.load_state_dict(c["state"])
        xo = c["x"].clone().requires_grad_(True)
        go = torch.autograd.grad(oc(xo, c["edge_index"], c["edge_weight"]).pow(2).sum(), [xo] + list(oc.parameters()))
        xg = c["x"].to(DEV).requires_grad_(True)
        gg = torch.autograd.grad(conv(xg, (c["edge_index"].to(DEV), c["edge_weight"].to(DEV))).pow(2).sum(),
                                 [xg] + list(conv.parameters()))
        for a, b in zip(gg, go):
            close(a, b)


def test_transfer_bit_exact_and_grads():
    gold = load_golden("neint.pt")
    tiny = gold["tiny"]
    ei = tiny["edge_index"].to(DEV)
    inc = Incidence(ei, 4)
    s2t = F_hl.edge_to_node(torch.tensor([[1.], [2.], [3.], [4.]], device=DEV), tiny["D"].to(DEV), inc)
    t2s = F_hl.node_to_edge(torch.tensor([[0.], [10.], [20.], [30.]], device=DEV), inc)
    assert torch.equal(s2t.cpu(), tiny["x_s2t"]) and torch.equal(t2s.cpu(), tiny["x_t2s"])
    b = make_batch("zinc", 32, seed=1)
    n, e = b.x_t.shape[0], b.x_s.shape[0]
    for width in (64, 192, 704, 6):
        x_t = torch.randn(n, width, requires_grad=True)
        x_s = torch.randn(e, width, requires_grad=True)
        D = O.degree(b.edge_index.view(-1), n) + 1e-6
        o_s2t, o_t2s = O.transfer(x_t, x_s, O.adj2par1(b.edge_index, n, e), D)
        inc = Incidence(b.edge_index.to(DEV), n)
        xt_g, xs_g = x_t.detach().to(DEV).requires_grad_(True), x_s.detach().to(DEV).requires_grad_(True)
        g_s2t = F_hl.edge_to_node(xs_g, D.to(DEV), inc)
        g_t2s = F_hl.node_to_edge(xt_g, inc)
        assert torch.equal(g_s2t.cpu(), o_s2t.detach()) and torch.equal(g_t2s.cpu(), o_t2s.detach())
        w1, w2 = torch.randn(n, width), torch.randn(e, width)
        go = torch.autograd.grad((o_s2t * w1).sum() + (o_t2s * w2).sum(), [x_t, x_s])
        gg = torch.autograd.grad((g_s2t * w1.to(DEV)).sum() + (g_t2s * w2.to(DEV)).sum(), [xt_g, xs_g])
        close(gg[0], go[0])
        close(gg[1], go[1])


@pytest.mark.parametrize("idx", range(3))
def test_node_edge_int_vs_golden_reference(idx):
    c = load_golden("neint.pt")["cases"][idx]
    n, e = c["x_t"].shape[0], c["x_s"].shape[0]
    sig = torch.nn.Sigmoid() if c["sigma"] == "sigmoid" else torch.nn.ReLU()
    mod = H.NodeEdgeInt(d=c["d"], dk=c["dk"], dv=c["dv"], only_att=c["only_att"], sigma=sig, l=c["l"]).to(DEV)
    mod.load_state_dict(c["state"], strict=False)
    mod.train()
    par = H.adj2par1(c["edge_index"].to(DEV), n, e)
    x_t, x_s = c["x_t"].to(DEV).requires_grad_(True), c["x_s"].to(DEV).requires_grad_(True)
    y_t, y_s = mod(x_t, x_s, par, c["D"].to(DEV))
    close(y_t, c["y_t"])
    close(y_s, c["y_s"])
    loss = (y_t * c["w_t"].to(DEV)).sum() + (y_s * c["w_s"].to(DEV)).sum()
    g = torch.autograd.grad(loss, [x_t, x_s] + list(mod.parameters()))
    close(g[0], c["gx_t"], atol=2e-5)
    close(g[1], c["gx_s"], atol=2e-5)
    for (nm, _), t in zip(mod.named_parameters(), g[2:]):
        close(t, c["gp"][nm], atol=2e-5)


def test_foreign_sparse_par_is_accepted():
    """NodeEdgeInt must also take a `par` built by the reference's own adj2par1 (plain sparse COO)."""
    c = load_golden("neint.pt")["cases"][0]
    n, e = c["x_t"].shape[0], c["x_s"].shape[0]
    par = O.adj2par1(c["edge_index"], n, e).to(DEV)
    mod = H.NodeEdgeInt(d=c["d"], dk=c["dk"], dv=c["dv"]).to(DEV)
    mod.load_state_dict(c["state"], strict=False)
    y_t, y_s = mod(c["x_t"].to(DEV), c["x_s"].to(DEV), par, c["D"].to(DEV))
    close(y_t, c["y_t"])
    close(y_s, c["y_s"])


def test_pool_block_vs_golden_reference():
    p = load_golden("pool.pt")
    n, e = p["x_t"].shape[0], p["x_s"].shape[0]
    pool = H.SAPool(d=6, dk=4).to(DEV)
    pool.load_state_dict(p["state"], strict=True)
    par = H.adj2par1(p["fine"]["edge_index"].to(DEV), n, e)
    coarse = SimpleNamespace(**{k: v for k, v in p["coarse"].items() if torch.is_tensor(v)})
    fine = SimpleNamespace()
    res = pool(p["x_t"].to(DEV), p["x_s"].to(DEV), par, p["D"].to(DEV), [fine, coarse],
               [p["c_node"].float().to(DEV)], [p["c_edge"].to(DEV)], 0, device=DEV)
    close(res[0], p["x_t1"])
    close(res[1], p["x_s1"])
    close(res[3], p["D1"])
    close(res[9], p["att_t"])
    close(res[10], p["att_s"])
    assert torch.equal(res[2].to_dense().cpu(), p["par1_dense"])


def test_segment_mean_and_gate_grads_vs_oracle():
    torch.manual_seed(0)
    r, f, ncl = 300, 48, 40
    src = torch.randn(r, f, requires_grad=True)
    att = torch.rand(r, 1, requires_grad=True)
    pos = torch.randint(0, ncl, (r, 1)).float()
    pos[::9] = float("inf")
    keep = ~torch.isinf(pos).view(-1)
    ref = O.scatter_mean((src * att)[keep], pos[keep].long())
    w = torch.randn_like(ref)
    go = torch.autograd.grad((ref * w).sum(), [src, att])
    sg, ag = src.detach().to(DEV).requires_grad_(True), att.detach().to(DEV).requires_grad_(True)
    out = F_hl.segment_mean(sg, F_hl.Segments.from_index(pos.to(DEV)), ag)
    assert torch.equal(out.cpu(), ref.detach())
    gg = torch.autograd.grad((out * w.to(DEV)).sum(), [sg, ag])
    close(gg[0], go[0])
    close(gg[1], go[1])
    # readout: contiguous segments
    counts = torch.tensor([5, 0, 17, 278])
    bvec = torch.repeat_interleave(torch.arange(4), counts)
    close(F_hl.segment_mean(sg, F_hl.Segments.from_counts(counts.to(DEV))), O.global_mean_pool(src, bvec, 4))
    # gate
    for sigma in ("sigmoid", "relu"):
        qc, qs, k = (torch.randn(r, 32, requires_grad=True) for _ in range(3))
        act = torch.sigmoid if sigma == "sigmoid" else torch.relu
        a_ref = act((0.3 * (qc * k).sum(1, keepdim=True) + 0.7 * (qs * k).sum(1, keepdim=True)) / 32 ** 0.5)
        wa = torch.randn(r, 1)
        go = torch.autograd.grad((a_ref * wa).sum(), [qc, qs, k])
        dq = [t.detach().to(DEV).requires_grad_(True) for t in (qc, qs, k)]
        a = F_hl.att_gate(dq[0], dq[1], dq[2], 0.7, sigma)
        close(a, a_ref)
        for x, y in zip(torch.autograd.grad((a * wa.to(DEV)).sum(), dq), go):
            close(x, y)


@pytest.mark.parametrize("shape", [(1000, 64), (3001, 256), (257, 10), (5, 7)])
@pytest.mark.parametrize("slope", [0.0, 0.1, 1.0])
def test_bn_act_vs_torch(shape, slope):
    torch.manual_seed(1)
    x = (torch.randn(*shape) * 3 + 5).requires_grad_(True)
    gamma, beta = torch.rand(shape[1], requires_grad=True), torch.randn(shape[1], requires_grad=True)
    z = torch.nn.functional.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-5)
    ref = torch.nn.functional.leaky_relu(z, slope) if slope != 1.0 else z
    w = torch.randn_like(ref)
    go = torch.autograd.grad((ref * w).sum(), [x, gamma, beta])
    xs = [t.detach().to(DEV).requires_grad_(True) for t in (x, gamma, beta)]
    y, stats = F_hl.bn_act_train(xs[0], xs[1], xs[2], 1e-5, slope)
    close(y, ref)
    close(stats[: shape[1]], x.mean(0))
    close(stats[shape[1]:], x.var(0, unbiased=False))
    for a, b in zip(torch.autograd.grad((y * w.to(DEV)).sum(), xs), go):
        close(a, b, atol=1e-4)


@pytest.mark.parametrize("K", [2, 3])
def test_zinc_model_vs_golden_reference(K):
    from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
    z = load_golden("zinc_model.pt")
    run = z["runs"][K]
    model = HL_HGCNN_zinc_dense_int3_pyr(K=K, **z["ctor"]).to(DEV)
    model.load_state_dict(run["state"], strict=True)
    model.train()
    data = SimpleNamespace(**{k: v.to(DEV) for k, v in z["batch"].items()})
    pred = model(data, device=DEV)
    close(pred, run["pred"], atol=2e-5)
    loss = torch.nn.functional.l1_loss(pred, data.y.view(-1, 1))
    g = torch.autograd.grad(loss, list(model.parameters()), allow_unused=True)
    for (n, _), t in zip(model.named_parameters(), g):
        ref = run["grads"][n]
        assert (t is None) == (ref is None), n
        if t is not None:
            close(t, ref, rtol=1e-3, atol=2e-5)      # 6-graph batch: BN over ~60 rows amplifies fp32 noise


def test_zinc_model_full_size_vs_oracle():
    """BASELINE config-1 model (filters 64/128/256, K=2) on a 64-graph ZINC-shaped batch against the CPU
    oracle.  Forward: every element of the prediction within rtol 1e-4 (the north-star bar).
    Gradients: each kernel meets 1e-4 on identical inputs (the per-op tests above); end to end they pass
    through 38 ReLU(BatchNorm(.)) layers whose masks flip for elements within rounding distance of zero, so
    even the fp32 CPU oracle is only ~2e-3 (up to 7e-3 on sign-cancelling bias gradients) away from its own
    fp64 run.  The end-to-end bar is therefore of that scale: 3e-2 per tensor, 1e-2 on average, vs fp64."""
    from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
    import copy
    torch.manual_seed(0)
    ctor = dict(channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[], K=2, node_dim=21, edge_dim=3, keig=7)
    ref = O.HL_HGCNN_zinc_dense_int3_pyr(**ctor)
    ref.train()
    b = make_batch("zinc", 64, seed=11)
    pred_ref = ref(b)
    ref64 = copy.deepcopy(ref).double()
    b64 = copy.copy(b)
    for k in ("x_t", "x_s", "y", "edge_weight_t", "edge_weight_s"):
        setattr(b64, k, getattr(b, k).double())
    pred64 = ref64(b64)
    g64 = torch.autograd.grad(torch.nn.functional.l1_loss(pred64, b64.y), list(ref64.parameters()), allow_unused=True)
    model = HL_HGCNN_zinc_dense_int3_pyr(**ctor).to(DEV)
    model.load_state_dict(ref.state_dict(), strict=True)
    model.train()
    d = batch_to(b, DEV)
    pred = model(d, device=DEV)
    close(pred, pred_ref, rtol=1e-4, atol=1e-4)
    close(pred, pred64.float(), rtol=1e-4, atol=1e-4)
    g = torch.autograd.grad(torch.nn.functional.l1_loss(pred, d.y), list(model.parameters()), allow_unused=True)
    rels = []
    for (n, _), a, r64 in zip(model.named_parameters(), g, g64):
        if r64 is None:
            assert a is None, n
            continue
        scale = float(r64.norm())
        if scale < 1e-6:                         # biases feeding a BatchNorm: exactly-zero true gradient
            assert float(a.norm()) < 1e-5, n
            continue
        rel = float((a.cpu().double() - r64).norm()) / scale
        assert rel < 3e-2, (n, rel)
        rels.append(rel)
    assert sum(rels) / len(rels) < 1e-2, sum(rels) / len(rels)
    print("gradient error vs fp64 oracle: mean", sum(rels) / len(rels), "max", max(rels))


def test_determinism_two_runs_bit_identical():
    from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
    torch.manual_seed(0)
    model = HL_HGCNN_zinc_dense_int3_pyr(channels=[1, 1], filters=[32, 64], K=3, node_dim=21, edge_dim=3, keig=7).to(DEV)
    d = batch_to(make_batch("zinc", 32, seed=2), DEV)
    outs = []
    for _ in range(2):
        model.zero_grad()
        p = model(d, device=DEV)
        p.abs().mean().backward()
        outs.append((p.detach().clone(), model.NEConv00.module_0.lins[1].weight.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("R,fo,fi", [(24000, 64, 64), (5000, 256, 704), (333, 5, 7), (1, 64, 28), (4097, 128, 10)])
def test_wgrad_and_colsum_vs_torch(R, fo, fi):
    torch.manual_seed(0)
    g, x = torch.randn(R, fo, device=DEV), torch.randn(R, fi + 8, device=DEV)[:, 4:4 + fi]
    ref = (g.double().t() @ x.double()).float()
    got = F_hl.wgrad(g, x)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-3 * max(1.0, R ** 0.5 / 30)), (got - ref).abs().max()
    wide = torch.zeros(fo, 2 * fi, device=DEV)
    F_hl.wgrad(g, x, wide[:, fi:])
    assert torch.equal(wide[:, fi:], got) and float(wide[:, :fi].abs().max()) == 0.0
    assert torch.equal(F_hl.wgrad(g, x), got)                                   # deterministic
    close(F_hl.colsum(g), g.double().sum(0).float(), atol=1e-3)


@pytest.mark.parametrize("width", [32, 64, 128, 256])
@pytest.mark.parametrize("K", [2, 4])
def test_staged_kernel_equals_per_row_kernel_fwd_and_bwd(width, K):
    """The row-window staged SpMM (cp.async.bulk ring) and the per-row kernel must agree bit for bit,
    forward basis and adjoint recurrence, on both operators of a batch large enough to take the staged path."""
    b = make_batch("zinc", 96, seed=21)
    n, e = b.x_t.shape[0], b.x_s.shape[0]
    ops = [CsrOperator(b.edge_index_t.to(DEV), b.edge_weight_t.to(DEV), n),
           CsrOperator(b.edge_index_s.to(DEV), b.edge_weight_s.to(DEV), e)]
    torch.manual_seed(width + K)
    xs = [torch.randn(n, width, device=DEV), torch.randn(e, width, device=DEV)]
    g0 = [torch.randn(n, width, device=DEV), torch.randn(e, width, device=DEV)]
    gt = [torch.randn(K - 1, n, width, device=DEV), torch.randn(K - 1, e, width, device=DEV)]
    res = {}
    try:
        for mode in (1, 0):
            N.lib().hl_set_spmm_mode(mode)
            for fam in (N.HL_LAGUERRE, N.HL_CHEB):
                t = F_hl.poly_basis_fwd(fam, K, ops, xs, width)
                a0, at = [t_.clone() for t_ in g0], [t_.clone() for t_ in gt]
                F_hl.poly_basis_bwd(fam, K, ops, a0, at, width)
                res[(mode, fam)] = (t, a0, at)
    finally:
        N.lib().hl_set_spmm_mode(0)
    for fam in (N.HL_LAGUERRE, N.HL_CHEB):
        for a, c in zip(res[(1, fam)], res[(0, fam)]):
            for u, v in zip(a, c):
                assert torch.equal(u, v), (fam, float((u - v).abs().max()))
    # and the adjoint really is the adjoint: with zero incoming gradient on T_0, the K=2 Laguerre pair gives
    # T_1 = x - A x and dx = G_1 - A^T G_1, so <T_1(x), G_1> == <x, dx> (in fp64 accumulation of fp32 results)
    for op, x, g in zip(ops, xs, gt):
        (t,) = F_hl.poly_basis_fwd(N.HL_LAGUERRE, 2, [op], [x], width)
        a0, a1 = torch.zeros_like(x), g[:1].clone()
        F_hl.poly_basis_bwd(N.HL_LAGUERRE, 2, [op], [a0], [a1], width)
        lhs = float((t[0].double() * g[0].double()).sum())
        rhs = float((x.double() * a0.double()).sum())
        assert abs(lhs - rhs) < 1e-5 * max(1.0, abs(lhs), float(t[0].double().norm() * g[0].double().norm()) * 1e-2), (lhs, rhs)


@pytest.mark.parametrize("M,Nn,K", [(300, 64, 64), (24001, 256, 704), (5000, 128, 28), (777, 32, 100), (1000, 512, 64),
                                     (1500, 704, 256), (1500, 448, 128), (900, 320, 64), (900, 192, 448)])
def test_tcgen05_dense_fp32_parity(M, Nn, K):
    """3xTF32 tensor-core GEMM: within rtol 1e-4 of fp64 (elementwise, scaled by the row/col magnitudes),
    for a @ w.T + bias, accumulate mode and the transposed-weight (data gradient) mode."""
    torch.manual_seed(M)
    a, w, bias = torch.randn(M, K, device=DEV), torch.randn(Nn, K, device=DEV), torch.randn(Nn, device=DEV)
    ref = a.double() @ w.double().t() + bias.double()
    scale = float(ref.abs().max())
    c = F_hl.dense(a, w, bias)
    assert float((c.double() - ref).abs().max()) < 1e-4 * scale
    c2 = F_hl.dense(a, w, None, out=c.clone(), accumulate=True)
    assert float((c2.double() - (2 * ref - bias.double())).abs().max()) < 2e-4 * scale
    g = torch.randn(M, Nn, device=DEV)
    if K % 16 == 0:
        d = F_hl.dense(g, w, transpose_w=True)
        refd = g.double() @ w.double()
        assert float((d.double() - refd).abs().max()) < 1e-4 * float(refd.abs().max())
    # a column-slice view as the A operand and as the weight (the split first MLP layer)
    wide = torch.randn(M, 2 * K, device=DEV)
    wcat = torch.randn(Nn, 2 * K, device=DEV)
    y = F_hl.dense(wide[:, :K], wcat[:, :K], bias)
    F_hl.dense(wide[:, K:], wcat[:, K:], None, out=y, accumulate=True)
    refy = wide.double() @ wcat.double().t() + bias.double()
    assert float((y.double() - refy).abs().max()) < 1e-4 * float(refy.abs().max())


@pytest.mark.parametrize("R,fo,fi", [(1472, 256, 448), (1600, 256, 704), (24000, 64, 64), (3000, 128, 192), (700, 256, 32),
                                     # 160 / 224 / 96-column tiles: 3-stage rings (odd: one converter group), long k-loops
                                     (60000, 64, 160), (60000, 128, 224), (40000, 64, 288), (230000, 32, 32)])
def test_tcgen05_wgrad_vs_fp64(R, fo, fi):
    torch.manual_seed(R)
    g, x = torch.randn(R, fo, device=DEV), torch.randn(R, fi, device=DEV)
    ref = g.double().t() @ x.double()
    got = F_hl.wgrad(g, x)
    assert float((got.double() - ref).abs().max()) < 1e-4 * float(ref.abs().max())
    assert torch.equal(F_hl.wgrad(g, x), got)


@pytest.mark.parametrize("R,fo,fi", [(1472, 256, 448), (24001, 64, 64), (3000, 128, 192), (5000, 12, 96), (300, 64, 64)])
def test_wgrad_with_folded_bias_gradient(R, fo, fi):
    """dbias = column sums of g from the same launches as dW (the converter warps that move g^T into tensor memory add
    up their column); shapes the tensor-core kernel does not take fall back to hl_wgrad + hl_colsum."""
    torch.manual_seed(R + fo)
    g, x = torch.randn(R, fo, device=DEV) + 0.3, torch.randn(R, fi, device=DEV)
    ref_w, ref_b = g.double().t() @ x.double(), g.double().sum(0)
    db = torch.full((fo,), float("nan"), device=DEV)
    dw = F_hl.wgrad(g, x, bias_out=db)
    assert float((dw.double() - ref_w).abs().max()) < 1e-4 * float(ref_w.abs().max())
    assert float((db.double() - ref_b).abs().max()) < 1e-5 * float(ref_b.abs().max())
    assert torch.equal(dw, F_hl.wgrad(g, x))                     # the fold does not change dW
    # accumulation into existing gradients (fused mode), weight as a column slice of a wider matrix
    wide = torch.randn(fo, fi + 32, device=DEV)
    acc_b = torch.randn(fo, device=DEV)
    want_w, want_b = wide[:, :fi].double() + ref_w, acc_b.double() + ref_b
    F_hl.wgrad(g, x, out=wide[:, :fi], accumulate=True, bias_out=acc_b, bias_accumulate=True)
    assert float((wide[:, :fi].double() - want_w).abs().max()) < 1e-4 * float(ref_w.abs().max())
    assert float((acc_b.double() - want_b).abs().max()) < 1e-5 * float(ref_b.abs().max())
    # deterministic
    db2 = torch.empty(fo, device=DEV)
    F_hl.wgrad(g, x, bias_out=db2)
    assert torch.equal(db, db2)


def test_bn_running_stats_match_torch():
    torch.manual_seed(3)
    x = torch.randn(777, 48) * 2 + 1
    ref = torch.nn.BatchNorm1d(48).train()
    with torch.no_grad():
        ref.running_mean.normal_()
        ref.running_var.uniform_(0.5, 2.0)
    mine = torch.nn.BatchNorm1d(48).to(DEV).train()
    mine.load_state_dict(ref.state_dict())
    from hlhgat_b200.lib.Hodge_Cheb_Conv import _bn_relu
    y = _bn_relu(mine, x.to(DEV), slope=1.0)
    close(y, ref(x))
    close(mine.running_mean, ref.running_mean)
    close(mine.running_var, ref.running_var)
    assert int(mine.num_batches_tracked) == 1


@pytest.mark.parametrize("R,fo,fi", [(24001, 64, 64), (24144, 256, 704), (5000, 128, 192), (3000, 256, 256), (700, 64, 96), (300, 64, 64)])
def test_wgrad2_two_operands_one_launch(R, fo, fi):
    """dW1 = g^T x1 and dW2 = g^T x2 from one launch + one reduce with two destinations (the two halves of the first
    NodeEdgeInt Linear read from strided dense-connection buffers; consecutive orders of a conv): fp64 parity, bias fold,
    accumulation, determinism; shapes the grouped kernel does not take fall back to two single launches."""
    torch.manual_seed(R + fi)
    g = torch.randn(R, fo, device=DEV) + 0.1
    wide = torch.randn(R, 2 * fi + 64, device=DEV)
    x1, x2 = wide[:, :fi], wide[:, fi + 32:2 * fi + 32]                # strided views, like the stack buffers
    r1, r2, rb = g.double().t() @ x1.double(), g.double().t() @ x2.double(), g.double().sum(0)
    w = torch.full((fo, 2 * fi), float("nan"), device=DEV)              # one weight matrix [fo, 2 fi]: the MLP case
    db = torch.full((fo,), float("nan"), device=DEV)
    F_hl.wgrad2(g, x1, x2, w[:, :fi], w[:, fi:], bias_out=db)
    scale = float(max(r1.abs().max(), r2.abs().max()))
    assert float((w[:, :fi].double() - r1).abs().max()) < 1e-4 * scale
    assert float((w[:, fi:].double() - r2).abs().max()) < 1e-4 * scale
    assert float((db.double() - rb).abs().max()) < 1e-5 * float(rb.abs().max())
    a1, a2 = torch.empty(fo, fi, device=DEV), torch.empty(fo, fi, device=DEV)   # two separate parameters: the conv case
    F_hl.wgrad2(g, x1, x2, a1, a2)
    assert torch.equal(a1, w[:, :fi]) and torch.equal(a2, w[:, fi:])             # deterministic, destination-independent
    acc1, acc2, accb = torch.randn(fo, fi, device=DEV), torch.randn(fo, fi, device=DEV), torch.randn(fo, device=DEV)
    want1, want2, wantb = acc1.double() + r1, acc2.double() + r2, accb.double() + rb
    F_hl.wgrad2(g, x1, x2, acc1, acc2, accumulate=True, bias_out=accb, bias_accumulate=True)
    assert float((acc1.double() - want1).abs().max()) < 1e-4 * scale and float((acc2.double() - want2).abs().max()) < 1e-4 * scale
    assert float((accb.double() - wantb).abs().max()) < 1e-5 * float(rb.abs().max())


# ------------------------------------------------------------------------------------------
# BatchNorm statistics from the GEMM epilogue (hl_gemm2_bn_tf32x3 + hl_bn_act_fwd_tiles)
def _merge_tiles(part, M, Nn, nvalid):
    """Chan merge of the per-32-row-block (mean | M2) pairs on the host in fp64 -> mean, biased variance."""
    nv = M if nvalid is None else min(int(nvalid), M)
    p = part.view(-1, 2, Nn).double().cpu()
    n, mean, m2 = 0.0, torch.zeros(Nn, dtype=torch.float64), torch.zeros(Nn, dtype=torch.float64)
    for k in range(p.shape[0]):
        cnt = max(0, min(32, nv - 32 * k))
        if cnt == 0:
            break
        d = p[k, 0] - mean
        tot = n + cnt
        mean = mean + d * cnt / tot
        m2 = m2 + p[k, 1] + d * d * n * cnt / tot
        n = tot
    return mean, m2 / max(n, 1.0)


@pytest.mark.parametrize("M,Nn,K,K2", [(300, 64, 64, 0), (24001, 256, 192, 64), (60, 64, 32, 0), (77000, 64, 128, 128),
                                       (5000, 128, 28, 0), (40000, 48, 64, 64), (1, 64, 64, 0), (129, 256, 704, 0)])
@pytest.mark.parametrize("nvalid", [None, 0.63])
def test_gemm_epilogue_block_statistics(M, Nn, K, K2, nvalid):
    g = torch.Generator().manual_seed(M + Nn)
    a1 = torch.randn(M, K, generator=g).to(DEV)
    w1 = (torch.randn(Nn, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(Nn, generator=g).to(DEV)
    nv_t = None if nvalid is None else torch.tensor([max(1, int(M * nvalid))], dtype=torch.int32, device=DEV)
    nv = None if nv_t is None else int(nv_t)
    req = F_hl.BnTiles(nv_t)
    if K2:
        a2 = (3.0 + torch.randn(M, K2, generator=g)).to(DEV)              # a mean far from zero: cancellation in E[x^2] - mean^2
        w2 = (torch.randn(Nn, K2, generator=g) / K2 ** 0.5).to(DEV)
        y = F_hl.dense2(a1, w1, a2, w2, bias, bn=req)
        y_plain = F_hl.dense2(a1, w1, a2, w2, bias)
    else:
        y = F_hl.dense(a1, w1, bias, bn=req)
        y_plain = F_hl.dense(a1, w1, bias)
    assert torch.equal(y, y_plain)                                        # the statistics do not change the output
    assert req.part is not None and req.matches(y, nv_t)
    mean, var = _merge_tiles(req.part, M, Nn, nv)
    yd = y[:nv].double().cpu() if nv is not None else y.double().cpu()
    close(mean.float(), yd.mean(0).float(), rtol=1e-5, atol=1e-6)
    close(var.float(), yd.var(0, unbiased=False).float(), rtol=1e-5, atol=1e-7)
    # accumulate launch: statistics of the FINAL values
    req2 = F_hl.BnTiles(nv_t)
    y2 = y.clone()
    F_hl.dense(a1, w1, None, out=y2, accumulate=True, bn=req2)
    mean2, var2 = _merge_tiles(req2.part, M, Nn, nv)
    y2d = y2[:nv].double().cpu() if nv is not None else y2.double().cpu()
    close(mean2.float(), y2d.mean(0).float(), rtol=1e-5, atol=1e-6)
    close(var2.float(), y2d.var(0, unbiased=False).float(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("M,Nn,K", [(24001, 64, 64), (3000, 256, 128), (77000, 128, 64), (50, 64, 64)])
@pytest.mark.parametrize("nvalid", [None, 0.8])
def test_bn_from_epilogue_tiles_equals_bn_with_own_statistics(M, Nn, K, nvalid):
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, K, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(Nn, K, generator=g) / K ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(Nn, generator=g).to(DEV).requires_grad_(True)
    gamma = torch.rand(Nn, generator=g).add(0.5).to(DEV).requires_grad_(True)
    beta = torch.randn(Nn, generator=g).to(DEV).requires_grad_(True)
    nv_t = None if nvalid is None else torch.tensor([int(M * nvalid)], dtype=torch.int32, device=DEV)
    dy = torch.randn(M, Nn, generator=g).to(DEV)
    res = []
    for use_tiles in (True, False):
        rm, rv = torch.zeros(Nn, device=DEV), torch.ones(Nn, device=DEV)
        cnt = torch.zeros((), dtype=torch.int64, device=DEV)
        with F_hl.bn_stats_from_epilogue(nv_t, enabled=use_tiles) as req:
            h = F_hl.linear(x, w, b)
        assert (req is not None and req.part is not None) == use_tiles
        y, stats = F_hl.bn_act_train(h, gamma, beta, 1e-5, 0.0, nv_t, rm, rv, 0.1, cnt, None, req)
        grads = torch.autograd.grad(y, (x, w, b, gamma, beta), dy)
        res.append((y, stats, rm, rv, *grads))
        assert int(cnt) == 1
    for i, (a, c) in enumerate(zip(*res)):                                # statistics differ in the last fp32 bit at most
        if i == 6:            # d(bias of the Linear): analytically zero behind a BatchNorm, both are cancellation noise
            assert float(a.abs().max()) < 1e-6 * M + 1e-5 and float(c.abs().max()) < 1e-6 * M + 1e-5
            continue
        close(a, c, rtol=1e-5, atol=1e-6 * max(1.0, float(c.detach().abs().max())))


def test_training_step_with_and_without_epilogue_statistics(monkeypatch):
    """Whole model, forward + backward: the epilogue statistics change nothing beyond fp32 rounding."""
    from hlhgat_b200.lib import Hodge_ST_Model as M
    from hlhgat_b200.workloads import WORKLOADS
    from hlhgat_b200.training import Capacity, pad_batch, StaticBatch
    wl = WORKLOADS["zinc"]
    raw = wl.make(64, 0)
    host = pad_batch(raw, Capacity.covering([raw]), deg_eps=wl.deg_eps)
    out = []
    for flag in (True, False):
        monkeypatch.setattr(F_hl, "_BN_EPILOGUE", flag)
        torch.manual_seed(0)
        model = getattr(M, wl.model)(**wl.ctor).to(DEV).train()
        batch = StaticBatch(host, torch.device(DEV))
        H.simplex.clear_caches()
        loss = wl.loss(model, batch)
        loss.backward()
        out.append((loss.detach(), [p.grad.clone() for p in model.parameters()]))
    close(out[0][0], out[1][0], rtol=1e-5)
    # gradients: the last-bit differences of the statistics pass through 38 BatchNorm + ReLU layers (mask flips of
    # near-zero pre-activations); same bar as the fp32-vs-fp64 model tests, measured here at ~5e-4 relative
    for ga, gb in zip(out[0][1], out[1][1]):
        assert float((ga - gb).norm()) <= 5e-3 * float(gb.norm()) + 1e-7


# ------------------------------------------------------------------------------------------
# weight gradients with their split reduces batched into one launch (hl_wgrad_deferred_tf32x3 + hl_wgrad_reduce_batch)
def test_deferred_weight_gradient_reduces_are_bit_identical(monkeypatch):
    monkeypatch.setattr(F_hl, "_WGRAD_DEFER", True)                       # opt-in path (HL_WGRAD_DEFER=1)
    g = torch.Generator().manual_seed(5)
    shapes = [(24001, 64, 64), (5000, 128, 192), (3000, 256, 256), (24144, 256, 704), (700, 64, 96), (26232, 64, 32)]
    shapes = shapes * 40                                                  # 240 descriptors: more than one parameter block
    cases = []
    for k, (R, fo, fi) in enumerate(shapes[:8]):
        cases.append((torch.randn(R, fo, generator=g).to(DEV), torch.randn(R, fi, generator=g).to(DEV), torch.randn(R, fi, generator=g).to(DEV)))

    def run(plan):
        outs = []
        ctx = plan if plan is not None else __import__("contextlib").nullcontext()
        with ctx:
            for k, (R, fo, fi) in enumerate(shapes):
                gg, x1, x2 = cases[k % len(cases)]
                if (gg.shape[0], gg.shape[1], x1.shape[1]) != (R, fo, fi):
                    gg, x1, x2 = [c for c in cases if (c[0].shape[0], c[0].shape[1], c[1].shape[1]) == (R, fo, fi)][0]
                w = torch.full((fo, 2 * fi), 0.5, device=DEV)
                b = torch.full((fo,), -1.0, device=DEV)
                if k % 3 == 0:
                    F_hl.wgrad2(gg, x1, x2, w[:, :fi], w[:, fi:], accumulate=True, bias_out=b, bias_accumulate=True)
                elif k % 3 == 1:
                    F_hl.wgrad(gg, x1, w[:, :fi], accumulate=True, bias_out=b, bias_accumulate=True)
                    F_hl.wgrad(gg, x2, w[:, fi:], accumulate=True)
                else:
                    F_hl.wgrad(gg, x1, w[:, :fi], accumulate=True)
                    F_hl.wgrad(gg, x2, w[:, :fi], accumulate=True)        # same destination twice: the second is not deferred
                outs.append((w, b))
            if plan is not None:
                assert len(plan.descs) > 192
        torch.cuda.synchronize()
        return outs

    ref = run(None)
    got = run(F_hl.WgradReducePlan())
    for k, ((w0, b0), (w1, b1)) in enumerate(zip(ref, got)):
        if k % 3 == 2:                          # (0.5 + a) + b against (0.5 + b) + a: the order of the two sums differs
            close(w0, w1, rtol=1e-6, atol=1e-4)
        else:
            assert torch.equal(w0, w1) and torch.equal(b0, b1)
    # and the values are right
    gg, x1, x2 = cases[1]
    k = 1
    w, b = got[k]
    fi = x1.shape[1]
    close(w[:, :fi], 0.5 + (gg.double().t() @ x1.double()).float(), rtol=1e-5, atol=1e-3)
    close(b, -1.0 + gg.double().sum(0).float(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M,Nn,K", [(38001, 64, 64), (40033, 48, 96), (20001, 416, 128), (9000, 800, 64), (38017, 128, 32),
                                    (12345, 672, 128), (300, 64, 64)])
@pytest.mark.parametrize("accumulate", [False, True])
def test_gemm_output_slice_is_written_exactly(M, Nn, K, accumulate):
    """The persistent kernel's TMA-store / TMA-reduce epilogue (and the register epilogue of the one-tile kernel) write the
    [M, N] slice of a wider, taller buffer and nothing around it: ragged row tiles, column tiles that are not multiples of
    32 (N = 416 -> 112-column tiles), N = 48."""
    g = torch.Generator().manual_seed(M + Nn)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(Nn, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(Nn, generator=g).to(DEV)
    big = torch.full((M + 70, Nn + 64), 7.25, device=DEV)
    out = big[3:3 + M, 32:32 + Nn]                              # 16-byte aligned start, row pitch N + 64
    if accumulate:
        out.copy_(torch.arange(M, device=DEV, dtype=torch.float32).remainder(5).unsqueeze(1).expand(M, Nn))
    old = out.clone()
    F_hl.dense(a, w, bias, out=out, accumulate=accumulate)
    ref = (a.double() @ w.double().t() + bias.double()).float() + (old if accumulate else 0.0)
    close(out, ref, rtol=1e-5, atol=2e-5)
    guard = big.clone()
    guard[3:3 + M, 32:32 + Nn] = 7.25
    assert bool((guard == 7.25).all())                          # nothing outside the slice was touched
