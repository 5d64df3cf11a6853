"""The drop-in boundary: the reference model classes' own call protocol --
gnn.Sequential blocks, int64 COO operators, sparse `par_1` from adj2par1, degree(), torch.cat dense connections, PyG
readout -- driving the hlhgat_b200 operator layer on CUDA, against the golden vectors of the UNMODIFIED reference.
(tests/reference_protocol.py restates the protocol; tests/test_oracle_golden.py pins it on the CPU; the real reference
sources are patched in tests/test_dropin_reference_sources.py, which only runs where /root/reference exists.)"""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

import hlhgat_b200 as H
from conftest import load_golden, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
import torch_geometric.nn as gnn  # noqa: E402  (shim: Sequential, BatchNorm, global_mean_pool -- the glue a PyG install provides)
from torch_geometric.utils import degree  # noqa: E402

from reference_protocol import ZincPyrProtocol, TspPyrProtocol  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OPS = SimpleNamespace(HodgeLaguerreConv=H.HodgeLaguerreConv, NodeEdgeInt=H.NodeEdgeInt, adj2par1=H.adj2par1)


def close(a, b, rtol=1e-4, atol=None):
    a, b = a.detach().cpu(), b.detach().cpu()
    atol = 1e-4 * float(b.abs().max()) if atol is None else atol
    assert a.shape == b.shape
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {(a - b).abs().max().item():.3e} (scale {b.abs().max():.3e})"


def to_dev(d):
    return SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in d.items()})


def check_grads(model, loss, ref_grads):
    g = torch.autograd.grad(loss, list(model.parameters()), allow_unused=True)
    for (n, _), t in zip(model.named_parameters(), g):
        ref = ref_grads[n]
        assert (t is None) == (ref is None), n
        if t is not None:                      # tiny batches: BN over ~40 rows amplifies fp32 noise
            close(t, ref, rtol=1e-3, atol=max(2e-5, 1e-4 * float(ref.abs().max())))


@pytest.mark.parametrize("K", [2, 3])
def test_zinc_reference_protocol_on_b200_ops_vs_golden(K):
    z = load_golden("zinc_model.pt")
    run = z["runs"][K]
    model = ZincPyrProtocol(OPS, gnn, degree, K=K, **z["ctor"]).to(DEV)
    model.load_state_dict(run["state"], strict=True)          # the reference checkpoint's keys, one for one
    model.train()
    data = to_dev(z["batch"])
    pred = model(data, device=DEV)
    close(pred, run["pred"], atol=2e-5)
    check_grads(model, torch.nn.functional.l1_loss(pred, data.y.view(-1, 1)), run["grads"])
    # the B200 operator layer really is what ran: the convs are hlhgat_b200 modules, the COO was bucketed once per batch
    assert type(model.HL_init_conv.module_0) is H.HodgeLaguerreConv
    assert len(data.edge_index_t._hl_ops) == 1 and len(data.edge_index_s._hl_ops) == 1


def test_tsp_reference_protocol_on_b200_ops_vs_golden():
    c = load_golden("models.pt")["tsp"]
    model = TspPyrProtocol(OPS, gnn, degree, **c["ctor"]).to(DEV)
    model.load_state_dict(c["state"], strict=True)
    model.train()
    pred, s_batch = model(to_dev(c["batch"]), device=DEV)
    close(pred, c["pred"], atol=2e-5)
    assert s_batch.shape[0] == pred.shape[0]
    check_grads(model, (pred * c["w"].to(DEV)).sum() / pred.shape[0], c["grads"])


def test_protocol_equals_mirrored_model_classes():
    """The mirrored model classes (hlhgat_b200.lib.Hodge_ST_Model: operators bucketed once, fused BatchNorm + ReLU kernels,
    no gnn.Sequential) and the reference protocol (torch BatchNorm1d / ReLU between the B200 operators) compute the same
    function on the same weights, to fp32 rounding."""
    from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
    from hlhgat_b200.synthetic import make_batch, batch_to
    torch.manual_seed(0)
    ctor = dict(channels=[1, 2], filters=[32, 64], mlp_channels=[], K=3, node_dim=21, edge_dim=3, keig=7)
    a = HL_HGCNN_zinc_dense_int3_pyr(**ctor).to(DEV).train()
    b = ZincPyrProtocol(OPS, gnn, degree, **ctor).to(DEV).train()
    b.load_state_dict(a.state_dict(), strict=True)
    d = batch_to(make_batch("zinc", 48, seed=5), DEV)
    close(a(d, device=DEV), b(d, device=DEV), rtol=1e-5, atol=1e-5)
