"""The TSP pyr, CIFAR10-superpixel attpool and peptides-func attpool callers on the GPU path, against
(i) golden vectors of the UNMODIFIED reference classes (tests/golden/make_golden_models.py) and
(ii) the CPU oracle on batches shaped like the BASELINE.json configs (rtol 1e-4 forward; gradients on
the end-to-end scale explained in test_gpu_parity.test_zinc_model_full_size_vs_oracle)."""
import copy
from types import SimpleNamespace

import pytest
import torch

import hlhgat_b200 as H
from hlhgat_b200 import functional as F_hl
from hlhgat_b200.lib import Hodge_ST_Model as M
from hlhgat_b200.simplex import incidence_for
from hlhgat_b200.synthetic import make_multilevel_batch, make_tsp_batch, make_batch
from oracle import hodge_oracle as O

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close(a, b, rtol=1e-4, atol=None):
    a, b = a.detach().cpu(), b.detach().cpu()
    atol = 1e-4 * float(b.abs().max()) if atol is None else atol
    assert a.shape == b.shape
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {(a - b).abs().max().item():.3e} (scale {b.abs().max():.3e})"


def to_dev(d):
    return SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in (d.items() if isinstance(d, dict) else vars(d).items())})


def check_grads_vs_golden(model, loss, ref_grads):
    g = torch.autograd.grad(loss, list(model.parameters()), allow_unused=True)
    for (n, _), t in zip(model.named_parameters(), g):
        ref = ref_grads[n]
        assert (t is None) == (ref is None), n
        if t is not None:                      # tiny batches: BN over ~40 rows amplifies fp32 noise
            close(t, ref, rtol=1e-3, atol=max(2e-5, 1e-4 * float(ref.abs().max())))


def test_boundary_absdiff_bit_exact_and_grad():
    torch.manual_seed(0)
    b = make_batch("cifar", 3, seed=2)
    n, e = b.x_t.shape[0], b.x_s.shape[0]
    for width in (1, 6, 64, 132):
        x = torch.randn(n, width)
        x[3] = x[b.edge_index[1][b.edge_index[0] == 3][0]]          # an exact tie: sgn(0) = 0 in the backward
        xo = x.clone().requires_grad_(True)
        par = O.adj2par1(b.edge_index, n, e)
        yo = torch.sparse.mm(par.transpose(0, 1), xo).abs() / 2
        w = torch.randn(e, width)
        (go,) = torch.autograd.grad((yo * w).sum(), xo)
        xg = x.to(DEV).requires_grad_(True)
        yg = F_hl.boundary_absdiff(xg, incidence_for(b.edge_index.to(DEV), n))
        (gg,) = torch.autograd.grad((yg * w.to(DEV)).sum(), xg)
        assert torch.equal(yg.cpu(), yo.detach())
        close(gg, go, rtol=1e-5, atol=1e-6)


def test_tsp_model_vs_golden_reference():
    c = load_golden("models.pt")["tsp"]
    model = M.HL_HGCNN_TSP_dense_int3_pyr(**c["ctor"]).to(DEV)
    model.load_state_dict(c["state"], strict=True)
    model.train()
    pred, s_batch = model(to_dev(c["batch"]), device=DEV)
    close(pred, c["pred"], atol=2e-5)
    assert s_batch.shape[0] == pred.shape[0] and s_batch.dtype == torch.int64
    check_grads_vs_golden(model, (pred * c["w"].to(DEV)).sum() / pred.shape[0], c["grads"])


@pytest.mark.parametrize("name", ["cifar", "pepfunc"])
def test_attpool_models_vs_golden_reference(name):
    c = load_golden("models.pt")[name]
    cls = {"cifar": M.HL_HGCNN_CIFAR10SP_dense_int3_attpool, "pepfunc": M.HL_HGCNN_pepfunc_dense_int3_attpool}[name]
    model = cls(**c["ctor"]).to(DEV)
    model.load_state_dict(c["state"], strict=True)
    model.train()
    datas = [to_dev(d) for d in c["datas"]]
    pred, att_t, att_s = model(datas, device=DEV, if_att=True)
    close(pred, c["pred"], atol=2e-5)
    close(att_t, c["att_t"], atol=1e-5)
    close(att_s, c["att_s"], atol=1e-5)
    check_grads_vs_golden(model, (pred * c["w"].to(DEV)).sum(), c["grads"])
    pred2 = model(datas, device=DEV)                         # default call: same prediction, no gate outputs
    close(pred2, pred, rtol=1e-6, atol=1e-6)


def _grad_report(model, g, g64):
    rels = []
    gmax = max(float(r.norm()) for r in g64 if r is not None)
    for (n, _), a, r64 in zip(model.named_parameters(), g, g64):
        if r64 is None:
            assert a is None, n
            continue
        scale = float(r64.norm())
        if scale < 1e-6 * max(1.0, gmax):          # biases feeding a BatchNorm: exactly-zero true gradient, fp32 noise
            assert a is None or float(a.norm()) < 1e-5 * max(1.0, gmax), n
            continue
        rel = float((a.cpu().double() - r64).norm()) / scale
        assert rel < 3e-2, (n, rel)
        rels.append(rel)
    assert sum(rels) / len(rels) < 1e-2
    return sum(rels) / len(rels), max(rels)


def _double(batch):
    out = copy.copy(batch)
    for k, v in vars(batch).items():
        if torch.is_tensor(v) and v.is_floating_point():
            setattr(out, k, v.double())
    return out


CONFIGS = {
    # BASELINE.json configs 3 and 4: the ctor arguments the reference scripts use
    "pepfunc": (M.HL_HGCNN_pepfunc_dense_int3_attpool, O.HL_HGCNN_pepfunc_dense_int3_attpool, "peptides",
                dict(channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[256], pool_loc=1, K=6, node_dim=9, edge_dim=3,
                     keig=10, num_classes=10), 6),
    "cifar": (M.HL_HGCNN_CIFAR10SP_dense_int3_attpool, O.HL_HGCNN_CIFAR10SP_dense_int3_attpool, "cifar",
              dict(channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[256], K=4, node_dim=5, edge_dim=4, keig=10,
                   pool_loc=1, l=0.5, num_classes=10), 6),
}


@pytest.mark.parametrize("name", ["pepfunc", "cifar"])
def test_attpool_models_config_size_vs_oracle(name):
    cls, ocls, shape, ctor, nb = CONFIGS[name]
    torch.manual_seed(0)
    ref = ocls(**ctor).train()
    datas = make_multilevel_batch(shape, nb, seed=3, node_dim=ctor["node_dim"] + ctor["keig"], edge_dim=ctor["edge_dim"] + ctor["keig"])
    w = torch.randn(nb, 10)
    pred_ref = ref(datas)
    ref64 = copy.deepcopy(ref).double()
    pred64 = ref64([_double(d) for d in datas])
    g64 = torch.autograd.grad((pred64 * w.double()).sum(), list(ref64.parameters()), allow_unused=True)
    model = cls(**ctor).to(DEV).train()
    model.load_state_dict(ref.state_dict(), strict=True)
    pred = model([to_dev(d) for d in datas], device=DEV)
    close(pred, pred_ref, rtol=1e-4, atol=1e-4 * float(pred_ref.abs().max()))
    close(pred, pred64.float(), rtol=1e-4, atol=1e-4 * float(pred_ref.abs().max()))
    g = torch.autograd.grad((pred * w.to(DEV)).sum(), list(model.parameters()), allow_unused=True)
    print(name, "gradient error vs fp64 oracle: mean %.2e max %.2e" % _grad_report(model, g, g64))


def test_tsp_model_config_size_vs_oracle():
    """BASELINE config 5 model (channels [4,4,4], filters [32,64,128], mlp [256], K=4) on two 120-node kNN-25
    graphs (the oracle's un-fused L1 propagate on full 500-node graphs takes minutes on the CPU)."""
    ctor = dict(channels=[4, 4, 4], filters=[32, 64, 128], mlp_channels=[256], K=4, node_dim=2, edge_dim=1, num_classes=2)
    torch.manual_seed(0)
    ref = O.HL_HGCNN_TSP_dense_int3_pyr(**ctor).train()
    b = make_tsp_batch(2, seed=4, n=120)
    w = torch.randn(b.x_s.shape[0], 2)
    pred_ref, _ = ref(b)
    ref64 = copy.deepcopy(ref).double()
    pred64, _ = ref64(_double(b))
    g64 = torch.autograd.grad((pred64 * w.double()).sum() / w.shape[0], list(ref64.parameters()), allow_unused=True)
    model = M.HL_HGCNN_TSP_dense_int3_pyr(**ctor).to(DEV).train()
    model.load_state_dict(ref.state_dict(), strict=True)
    pred, s_batch = model(to_dev(b), device=DEV)
    close(pred, pred_ref, rtol=1e-4, atol=1e-4 * float(pred_ref.abs().max()))
    close(pred, pred64.float(), rtol=1e-4, atol=1e-4 * float(pred_ref.abs().max()))
    g = torch.autograd.grad((pred * w.to(DEV)).sum() / w.shape[0], list(model.parameters()), allow_unused=True)
    print("tsp gradient error vs fp64 oracle: mean %.2e max %.2e" % _grad_report(model, g, g64))


# ---------------------------------------------------------------------------------------------
# factored Hodge 1-Laplacian (opt-in): L1 x = diag(s) B1^T (B1 x)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,batch,width", [("cifar", 6, 64), ("tsp", 2, 32), ("zinc", 40, 128), ("peptides", 5, 20), ("cifar", 3, 7)])
@pytest.mark.parametrize("K", [2, 5])
def test_factored_hodge1_basis_equals_csr_path(shape, batch, width, K):
    from hlhgat_b200 import _native as N
    from hlhgat_b200.simplex import CsrOperator, Hodge1Factor
    b = make_batch(shape, batch, seed=13)
    n, e = b.x_t.shape[0], b.x_s.shape[0]
    op = CsrOperator(b.edge_index_s.to(DEV), b.edge_weight_s.to(DEV), e)
    op.factored = Hodge1Factor.from_operator(op, incidence_for(b.edge_index.to(DEV), n))
    torch.manual_seed(K)
    x = torch.randn(e, width, device=DEV)
    g0, gt = torch.randn(e, width, device=DEV), torch.randn(K - 1, e, width, device=DEV)
    res = {}
    try:
        for flag in (False, True):
            F_hl.enable_factored_hodge1(flag)
            for fam in (N.HL_LAGUERRE, N.HL_CHEB):
                (t,) = F_hl.poly_basis_fwd(fam, K, [op], [x], width)
                a0, at = g0.clone(), gt.clone()
                F_hl.poly_basis_bwd(fam, K, [op], [a0], [at], width)
                res[(flag, fam)] = (t, a0, at)
    finally:
        F_hl.enable_factored_hodge1(False)
    for fam in (N.HL_LAGUERRE, N.HL_CHEB):
        for u, v in zip(res[(True, fam)], res[(False, fam)]):
            close(u, v, rtol=1e-5, atol=2e-5 * float(v.abs().max()))


def test_factored_hodge1_from_constructor_and_in_model():
    """The GPU constructor attaches the factor itself; a TSP model step with the option on equals the CSR step."""
    from hlhgat_b200.construct import build_simplex_batch
    b = make_tsp_batch(2, seed=6, n=80, k=8)
    ei = b.edge_index
    sb = build_simplex_batch(torch.cat([ei[0], ei[1]]).to(DEV), torch.cat([ei[1], ei[0]]).to(DEV), b.num_node1)
    ref = sb.op_s.fwd[2]
    from hlhgat_b200.simplex import Hodge1Factor
    derived = Hodge1Factor.from_operator(sb.op_s, sb.incidence)
    close(sb.op_s.factored.edge_scale, derived.edge_scale, rtol=1e-6, atol=0)
    assert ref.numel() > 0
    ctor = dict(channels=[1, 1], filters=[32, 64], mlp_channels=[32], K=4, node_dim=2, edge_dim=1, num_classes=1)
    torch.manual_seed(0)
    model = M.HL_HGCNN_TSP_dense_int3_pyr(**ctor).to(DEV).train()
    d = to_dev(b)
    w = torch.randn(b.x_s.shape[0], 1, device=DEV)
    outs = []
    try:
        for flag in (False, True):
            F_hl.enable_factored_hodge1(flag)
            from hlhgat_b200.simplex import clear_caches
            clear_caches()
            pred, _ = model(d, device=DEV)
            g = torch.autograd.grad((pred * w).sum(), [p for p in model.parameters()], allow_unused=True)
            outs.append((pred, g))
    finally:
        F_hl.enable_factored_hodge1(False)
    close(outs[1][0], outs[0][0], rtol=1e-4, atol=1e-4 * float(outs[0][0].abs().max()))
    gmax = max(float(c.norm()) for c in outs[0][1] if c is not None)
    for a, c in zip(outs[1][1], outs[0][1]):
        if c is not None and float(c.norm()) > 1e-4 * gmax:        # (biases in front of a BatchNorm: pure rounding noise)
            assert float((a - c).norm()) < 2e-3 * float(c.norm())
