"""Artefacts the reference ships as known answers: the state_dict key / shape contract of its only
checkpoint (HL-HGAT-DEMO/weights/HL_HGAT_Brain.pt -> tests/golden/brain_ckpt_contract.json) on the CPU, and the DEMO's
real brain skeleton (Group_FC.mat / Group_FCMask.mat -> tests/golden/group_fc.pt) through the GPU constructor."""
import json
import os

import pytest
import torch
import torch.nn as nn

from conftest import GOLDEN, load_golden


def _brain_modules():
    """The operator-layer part of the checkpointed model (filters 32 / 64 / 128, two layers per stage, K = 4, one
    attention gate on the 96-wide dense-connection buffer, a two-conv K = 1 readout), built from hlhgat_b200 modules
    alone.  node_embedding.* (Inception1D over the fMRI series) and the final Linear are outside the path."""
    import hlhgat_b200 as H
    from hlhgat_b200.lib.Hodge_ST_Model import _SeqConv  # noqa: F401
    m = nn.Module()
    K = 4
    m.HL_init_conv = H.NEConv(64, 1, 32, K)
    fin = 32
    for i, f in enumerate([32, 64, 128]):
        for j in range(2):
            setattr(m, f"NEInt{i}{j}", H.NodeEdgeInt(d=fin, dv=f))
            setattr(m, f"NEConv{i}{j}", H.NEConv(f, f, f, K))
            fin += f
        if i == 0:
            m.NEAtt0 = H.NodeEdgeInt(d=fin, dk=32, only_att=True)
    ro = nn.Module()
    ro.module_0 = H.HodgeLaguerreConv(128, 1, K=1)
    ro.module_1 = H.HodgeLaguerreConv(128, 1, K=1)
    m.readout = ro
    return m


def test_brain_checkpoint_key_and_shape_contract():
    contract = json.load(open(os.path.join(GOLDEN, "brain_ckpt_contract.json")))
    on_path = {k: tuple(v) for k, v in contract.items() if not k.startswith(("node_embedding.", "out."))}
    mine = {k: tuple(v.shape) for k, v in _brain_modules().state_dict().items()}
    assert set(mine) == set(on_path), (sorted(set(mine) ^ set(on_path))[:10])
    for k, shp in on_path.items():
        assert mine[k] == shp, (k, mine[k], shp)
    # ... so the checkpoint's tensors load one for one
    fake = {k: torch.zeros(s, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32) for k, s in on_path.items()}
    _brain_modules().load_state_dict(fake, strict=True)


@pytest.mark.gpu
def test_group_fc_known_answers_through_gpu_constructor():
    """N = 268, E = 8,997, lambda_max = 143.2766, nnz(L0) = N + 2E = 18,262, nnz(L1) = 1,371,129, max degree 142 -- a
    real, high-degree graph (rows of L1 up to 283 nonzeros) through hl_build_edges / hl_lambda_max / hl_laplacian_*."""
    from hlhgat_b200.construct import build_simplex_batch
    g = load_golden("group_fc.pt")
    dev = "cuda:0"
    und = g["edge_index"].long()
    src, dst = torch.cat([und[0], und[1]]), torch.cat([und[1], und[0]])           # both directions, as a dataset delivers them
    perm = torch.randperm(src.numel(), generator=torch.Generator().manual_seed(0))
    sb = build_simplex_batch(src[perm].to(dev), dst[perm].to(dev), torch.tensor([g["num_nodes"]]))
    assert (sb.num_nodes, sb.num_edges) == (268, 8997) == (g["num_nodes"], g["num_edges"])
    assert torch.equal(sb.edge_index.cpu(), und)                                   # lexicographic i < j order
    lam = float(sb.lambda_max[0])
    assert abs(lam - float(g["maxeig64"])) < 2e-6 * float(g["maxeig64"])          # fp64 Lanczos vs fp64 eigh
    assert abs(lam - float(g["maxeig"])) < 1e-5 * float(g["maxeig"])              # the reference's own fp32 eigh: 143.2766
    ei_t, w_t = sb.coo("t")
    ei_s, w_s = sb.coo("s")
    assert ei_t.shape[1] == g["nnz_t"] == 18262 and ei_s.shape[1] == g["nnz_s"] == 1371129
    assert int(sb.incidence.degree().max()) == g["max_degree"] == 142
    assert torch.equal(ei_t.cpu(), g["ei_t"].long())
    assert torch.allclose(w_t.cpu(), g["w_t"], rtol=1e-5, atol=0)
    assert torch.equal(ei_s[:, g["pick"].to(dev)].cpu(), g["pick_ei_s"].long())    # 4096 entries spread over the 1.37 M
    assert torch.allclose(w_s[g["pick"].to(dev)].cpu(), g["pick_w_s"], rtol=1e-5, atol=0)
    assert abs(float(w_s.double().abs().sum()) - g["sum_abs_w_s"]) < 1e-5 * g["sum_abs_w_s"]
    assert abs(float(w_t.double().sum()) - g["sum_w_t"]) < 1e-4
    diag = w_s[ei_s[0] == ei_s[1]]
    assert diag.numel() == 8997 and torch.allclose(diag, torch.full_like(diag, 4.0 / lam), rtol=1e-6)
    # the polynomial SpMM on this operator (long rows) against a dense product with the same matrix
    from hlhgat_b200 import functional as F_hl, _native as N
    torch.manual_seed(0)
    x = torch.randn(8997, 32, device=dev)
    (t,) = F_hl.poly_basis_fwd(N.HL_LAGUERRE, 2, [sb.op_s], [x], 32)
    dense = torch.zeros(8997, 8997, device=dev, dtype=torch.float64)
    dense.index_put_((ei_s[1], ei_s[0]), w_s.double(), accumulate=True)
    ref = x.double() - dense @ x.double()
    assert float((t[0].double() - ref).abs().max()) < 1e-5 * float(ref.abs().max())
