"""The oracle (oracle/hodge_oracle.py) against the golden vectors produced by the unmodified
reference modules (tests/golden/make_golden.py) and the analytic known answers of SURVEY.md
section 4.  CPU only."""
import math
from types import SimpleNamespace

import pytest
import torch

from oracle import hodge_oracle as O

from conftest import load_golden

TOL = dict(rtol=1e-5, atol=1e-6)


def _conv(case):
    cls = O.HodgeLaguerreConv if case["family"] == "laguerre" else O.HodgeChebConv
    conv = cls(case["fin"], case["fout"], case["K"])
    conv.load_state_dict(case["state"], strict=True)
    return conv


@pytest.mark.parametrize("idx", range(21))
def test_conv_matches_reference(idx):
    cases = load_golden("conv.pt")["cases"]
    assert len(cases) == 21
    c = cases[idx]
    conv = _conv(c)
    x = c["x"].clone().requires_grad_(True)
    y = conv(x, c["edge_index"], c["edge_weight"])
    assert torch.equal(y, c["y"]) or torch.allclose(y, c["y"], **TOL)
    g = torch.autograd.grad((y * c["wsum"]).sum(), [x] + list(conv.parameters()))
    assert torch.allclose(g[0], c["gx"], **TOL)
    for (n, _), t in zip(conv.named_parameters(), g[1:]):
        assert torch.allclose(t, c["gp"][n], **TOL), n


def test_fastconv_quirk_matches_demo():
    for c in load_golden("fastconv.pt"):
        conv = O.HodgeLaguerreFastConv(5, 6, c["K"], quirk=True)
        conv.load_state_dict(c["state"], strict=True)
        assert torch.allclose(conv(c["x"], c["edge_index"], c["edge_weight"]), c["y"], **TOL)
        if c["K"] >= 3:
            fixed = O.HodgeLaguerreFastConv(5, 6, c["K"], quirk=False)
            fixed.load_state_dict(c["state"])
            assert not torch.allclose(fixed(c["x"], c["edge_index"], c["edge_weight"]), c["y"], **TOL)


@pytest.mark.parametrize("idx", range(3))
def test_node_edge_int_matches_reference(idx):
    gold = load_golden("neint.pt")
    c = gold["cases"][idx]
    n, e = c["x_t"].shape[0], c["x_s"].shape[0]
    sig = torch.nn.Sigmoid() if c["sigma"] == "sigmoid" else torch.nn.ReLU()
    mod = O.NodeEdgeInt(d=c["d"], dk=c["dk"], dv=c["dv"], only_att=c["only_att"], sigma=sig, l=c["l"])
    mod.load_state_dict(c["state"], strict=False)
    mod.train()
    par = O.adj2par1(c["edge_index"], n, e)
    x_t, x_s = c["x_t"].clone().requires_grad_(True), c["x_s"].clone().requires_grad_(True)
    y_t, y_s = mod(x_t, x_s, par, c["D"])
    assert torch.allclose(y_t, c["y_t"], **TOL) and torch.allclose(y_s, c["y_s"], **TOL)
    loss = (y_t * c["w_t"]).sum() + (y_s * c["w_s"]).sum()
    g = torch.autograd.grad(loss, [x_t, x_s] + list(mod.parameters()))
    assert torch.allclose(g[0], c["gx_t"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(g[1], c["gx_s"], rtol=1e-4, atol=1e-5)
    for (nm, _), t in zip(mod.named_parameters(), g[2:]):
        assert torch.allclose(t, c["gp"][nm], rtol=1e-4, atol=1e-5), nm


def test_tiny_graph_known_answers():
    """SURVEY.md section 4 'tiny golden graph' + the reference's own numbers for it."""
    tiny = load_golden("neint.pt")["tiny"]
    ei = torch.tensor([[0, 0, 1, 2], [1, 2, 2, 3]])
    par = O.adj2par1(ei, 4, 4)
    assert torch.equal(par.to_dense(), tiny["par_dense"])
    assert torch.equal(par.to_dense(), torch.tensor(
        [[-1., -1, 0, 0], [1, 0, -1, 0], [0, 1, 1, -1], [0, 0, 0, 1]]))
    g = O.build_simplex_graph(torch.cat([ei, ei.flip(0)], 1), 4)
    assert abs(float(g.maxeig) - 4.0) < 1e-5
    assert g.edge_index_t.tolist() == [[0, 0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 3], [0, 1, 2, 0, 1, 2, 0, 1, 2, 3, 2, 3]]
    assert torch.allclose(g.edge_weight_t, torch.tensor([1, -.5, -.5, -.5, 1, -.5, -.5, -.5, 1.5, -.5, -.5, .5]), atol=1e-6)
    assert g.edge_index_s.tolist() == [[0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3], [0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3]]
    assert torch.allclose(g.edge_weight_s, torch.tensor([1, .5, -.5, .5, 1, .5, -.5, -.5, .5, 1, -.5, -.5, -.5, 1]), atol=1e-6)
    D = O.degree(ei.view(-1), 4) + 1e-6
    x_s = torch.tensor([[1.], [2.], [3.], [4.]])
    x_t = torch.tensor([[0.], [10.], [20.], [30.]])
    s2t, t2s = O.transfer(x_t, x_s, par, D)
    assert torch.equal(s2t, tiny["x_s2t"]) and torch.equal(t2s, tiny["x_t2s"])
    assert torch.allclose(s2t.view(-1), torch.tensor([1.5, 2., 3., 4.]), rtol=2e-6)
    assert t2s.view(-1).tolist() == [5., 10., 15., 25.]


def test_construction_matches_reference_bit_exact():
    for c in load_golden("construct.pt"):
        g = O.build_simplex_graph(c["ei_dir"], c["n"])
        assert torch.equal(g.edge_index, c["edge_index"])
        assert torch.equal(g.edge_index_t, c["edge_index_t"])
        assert torch.equal(g.edge_index_s, c["edge_index_s"])
        assert torch.equal(g.edge_weight_t, c["edge_weight_t"])
        assert torch.equal(g.edge_weight_s, c["edge_weight_s"])
        assert torch.equal(g.maxeig, c["maxeig"])


def test_pool_block_matches_reference():
    p = load_golden("pool.pt")
    n, e = p["x_t"].shape[0], p["x_s"].shape[0]
    att = O.NodeEdgeInt(d=6, dk=4, only_att=True)
    att.load_state_dict({k.replace("NEAtt.", ""): v for k, v in p["state"].items()}, strict=True)
    par = O.adj2par1(p["fine"]["edge_index"], n, e)
    a_t, a_s = att(p["x_t"], p["x_s"], par, p["D"])
    assert torch.allclose(a_t, p["att_t"], **TOL) and torch.allclose(a_s, p["att_s"], **TOL)
    x_t1, x_s1 = O.attention_pool(p["x_t"], p["x_s"], a_t, a_s, p["c_node"].float(), p["c_edge"])
    assert torch.allclose(x_t1, p["x_t1"], **TOL) and torch.allclose(x_s1, p["x_s1"], **TOL)
    par1 = O.adj2par1(p["coarse"]["edge_index"], x_t1.shape[0], x_s1.shape[0])
    assert torch.equal(par1.to_dense(), p["par1_dense"])
    assert torch.equal(O.degree(p["coarse"]["edge_index"].view(-1), x_t1.shape[0]) + 1e-6, p["D1"])


@pytest.mark.parametrize("K", [2, 3])
def test_zinc_model_matches_reference(K):
    z = load_golden("zinc_model.pt")
    run = z["runs"][K]
    model = O.HL_HGCNN_zinc_dense_int3_pyr(K=K, **z["ctor"])
    model.load_state_dict(run["state"], strict=True)      # same key names as the reference
    model.train()
    data = SimpleNamespace(**z["batch"])
    pred = model(data)
    assert torch.allclose(pred, run["pred"], rtol=1e-5, atol=1e-5)
    loss = torch.nn.functional.l1_loss(pred, data.y.view(-1, 1))
    g = torch.autograd.grad(loss, list(model.parameters()), allow_unused=True)
    for (n, _), t in zip(model.named_parameters(), g):
        ref = run["grads"][n]
        assert (t is None) == (ref is None), n
        if t is not None:
            assert torch.allclose(t, ref, rtol=1e-3, atol=max(2e-5, 1e-4 * float(ref.abs().max()))), n


def test_collate_matches_reference_batch():
    z = load_golden("zinc_model.pt")
    graphs = []
    off_n = off_e = 0
    b = z["batch"]
    for g in z["raw"]:
        c = O.build_simplex_graph(g["ei_dir"], g["n"])
        e = c.edge_index.shape[1]
        c.x_t = b["x_t"][off_n:off_n + g["n"]]
        c.x_s = b["x_s"][off_e:off_e + e]
        c.y = None
        off_n, off_e = off_n + g["n"], off_e + e
        graphs.append(c)
    bb = O.collate(graphs)
    for k in ("edge_index", "edge_index_t", "edge_index_s", "edge_weight_t", "edge_weight_s", "num_node1", "num_edge1"):
        assert torch.equal(getattr(bb, k), b[k]), k


@pytest.mark.parametrize("family", ["laguerre", "cheb"])
def test_scalar_polynomial_identity(family):
    """Diagonal operator with weights w and W_k = delta_kj: the conv returns L_j(w) / T_j(w)."""
    w = torch.linspace(0.0, 2.0, 9, dtype=torch.float64)
    ei = torch.arange(9).repeat(2, 1)
    x = torch.ones(9, 1, dtype=torch.float64)
    polys = {"laguerre": [lambda t: t * 0 + 1, lambda t: 1 - t, lambda t: (t * t - 4 * t + 2) / 2,
                          lambda t: (-t ** 3 + 9 * t * t - 18 * t + 6) / 6],
             "cheb": [lambda t: t * 0 + 1, lambda t: t, lambda t: 2 * t * t - 1, lambda t: 4 * t ** 3 - 3 * t]}[family]
    cls = O.HodgeLaguerreConv if family == "laguerre" else O.HodgeChebConv
    for j in range(4):
        conv = cls(1, 1, 4).double()
        with torch.no_grad():
            for k, lin in enumerate(conv.lins):
                lin.weight.fill_(1.0 if k == j else 0.0)
        assert torch.allclose(conv(x, ei, w).view(-1), polys[j](w), atol=1e-12)


def test_gradcheck_fp64():
    c = load_golden("conv.pt")["cases"][2 * 3 + 1]          # laguerre K=3, edge side
    conv = _conv(c).double()
    x = c["x"].double().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda v: conv(v, c["edge_index"], c["edge_weight"].double()), (x,))


# --------------------------------------------------------------------------------------
# the TSP / CIFAR10-superpixel / peptides-func callers (tests/golden/make_golden_models.py)
# --------------------------------------------------------------------------------------
def _check_grads(model, loss, ref_grads):
    g = torch.autograd.grad(loss, list(model.parameters()), allow_unused=True)
    for (n, _), t in zip(model.named_parameters(), g):
        ref = ref_grads[n]
        assert (t is None) == (ref is None), n
        if t is not None:            # (a bias in front of a BatchNorm has a zero true gradient: rounding noise up to ~7e-6)
            assert torch.allclose(t, ref, rtol=1e-4, atol=max(1e-5, 2e-5 * float(ref.abs().max()))), n


def test_tsp_model_matches_reference():
    c = load_golden("models.pt")["tsp"]
    model = O.HL_HGCNN_TSP_dense_int3_pyr(**c["ctor"])
    model.load_state_dict(c["state"], strict=True)
    model.train()
    pred, s_batch = model(SimpleNamespace(**c["batch"]))
    assert torch.allclose(pred, c["pred"], rtol=1e-5, atol=1e-5)
    assert s_batch.numel() == pred.shape[0]
    _check_grads(model, (pred * c["w"]).sum() / pred.shape[0], c["grads"])


@pytest.mark.parametrize("name", ["cifar", "pepfunc"])
def test_attpool_models_match_reference(name):
    c = load_golden("models.pt")[name]
    cls = {"cifar": O.HL_HGCNN_CIFAR10SP_dense_int3_attpool, "pepfunc": O.HL_HGCNN_pepfunc_dense_int3_attpool}[name]
    model = cls(**c["ctor"])
    model.load_state_dict(c["state"], strict=True)
    model.train()
    datas = [SimpleNamespace(**d) for d in c["datas"]]
    pred, att_t, att_s = model(datas, if_att=True)
    assert torch.allclose(pred, c["pred"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(att_t, c["att_t"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(att_s, c["att_s"], rtol=1e-5, atol=1e-6)
    _check_grads(model, (pred * c["w"]).sum(), c["grads"])


# ---------------------------------------------------------------------------------------------------------------
# tests/reference_protocol.py (the reference model classes' call protocol with an injected operator layer) pinned to
# the unmodified reference: with the ORACLE operator layer on the CPU it must reproduce the golden predictions and
# gradients.  tests/test_gpu_dropin.py runs the same classes on the hlhgat_b200 operator layer.
# ---------------------------------------------------------------------------------------------------------------
def _protocol_env():
    import os
    import sys
    from conftest import ROOT
    shim = os.path.join(ROOT, "oracle", "pyg_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import torch_geometric.nn as gnn
    from torch_geometric.utils import degree
    ops = SimpleNamespace(HodgeLaguerreConv=O.HodgeLaguerreConv, NodeEdgeInt=O.NodeEdgeInt, adj2par1=O.adj2par1)
    return ops, gnn, degree


@pytest.mark.parametrize("K", [2, 3])
def test_reference_protocol_zinc_pinned_to_golden(K):
    from reference_protocol import ZincPyrProtocol
    ops, gnn, degree = _protocol_env()
    z = load_golden("zinc_model.pt")
    run = z["runs"][K]
    model = ZincPyrProtocol(ops, gnn, degree, K=K, **z["ctor"])
    model.load_state_dict(run["state"], strict=True)
    model.train()
    data = SimpleNamespace(**z["batch"])
    pred = model(data, device="cpu")
    assert torch.allclose(pred, run["pred"], rtol=1e-5, atol=1e-6)
    g = torch.autograd.grad(torch.nn.functional.l1_loss(pred, data.y.view(-1, 1)), list(model.parameters()), allow_unused=True)
    for (n, _), t in zip(model.named_parameters(), g):
        ref = run["grads"][n]
        assert (t is None) == (ref is None), n
        if t is not None:
            assert torch.allclose(t, ref, rtol=1e-3, atol=max(2e-5, 1e-4 * float(ref.abs().max()))), n


def test_reference_protocol_tsp_pinned_to_golden():
    from reference_protocol import TspPyrProtocol
    ops, gnn, degree = _protocol_env()
    c = load_golden("models.pt")["tsp"]
    model = TspPyrProtocol(ops, gnn, degree, **c["ctor"])
    model.load_state_dict(c["state"], strict=True)
    model.train()
    pred, _ = model(SimpleNamespace(**c["batch"]), device="cpu")
    assert torch.allclose(pred, c["pred"], rtol=1e-4, atol=1e-5), (pred - c["pred"]).abs().max()
    g = torch.autograd.grad((pred * c["w"]).sum() / pred.shape[0], list(model.parameters()), allow_unused=True)
    for (n, _), t in zip(model.named_parameters(), g):
        ref = c["grads"][n]
        assert (t is None) == (ref is None), n
        if t is not None:
            assert torch.allclose(t, ref, rtol=1e-3, atol=max(2e-5, 1e-4 * float(ref.abs().max()))), n
