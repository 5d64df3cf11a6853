"""Padded fixed-capacity batches and whole-step CUDA-graph replay give the same training step as
the plain eager path on the unpadded batch."""
import copy

import pytest
import torch

import hlhgat_b200  # noqa: F401
from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
from hlhgat_b200.parallel import FlatGradBucket
from hlhgat_b200.synthetic import make_batch, batch_to
from hlhgat_b200.training import Capacity, pad_batch, GraphedTrainStep, StaticBatch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CTOR = dict(channels=[1, 2], filters=[32, 64], mlp_channels=[48], K=3, node_dim=21, edge_dim=3, keig=7)


def test_padded_batch_equals_unpadded():
    torch.manual_seed(0)
    raws = [make_batch("zinc", 48, seed=s) for s in (1, 2)]
    cap = Capacity.covering(raws, slack=0.1)
    model = HL_HGCNN_zinc_dense_int3_pyr(**CTOR).to(DEV).train()
    ref = copy.deepcopy(model)
    b = batch_to(raws[0], DEV)
    pred = ref(b, device=DEV)
    loss = torch.nn.functional.l1_loss(pred, b.y)
    g_ref = torch.autograd.grad(loss, list(ref.parameters()))
    sb = StaticBatch(pad_batch(raws[0], cap), DEV)
    pred_p = model(sb, device=DEV)
    assert pred_p.shape[0] == 49                         # 48 graphs + the ghost graph
    assert torch.allclose(pred_p[:48], pred, rtol=1e-4, atol=1e-5)
    loss_p = torch.nn.functional.l1_loss(pred_p[:48], sb.y)
    g = torch.autograd.grad(loss_p, list(model.parameters()))
    for (n, _), a, r in zip(model.named_parameters(), g, g_ref):
        assert float((a - r).norm()) < 1e-3 * float(r.norm()) + 1e-7 * r.numel() ** 0.5, n
    # running statistics must ignore the ghost rows as well
    for (n, a), (_, r) in zip(model.named_buffers(), ref.named_buffers()):
        assert torch.allclose(a.float(), r.float(), rtol=1e-4, atol=1e-6), n


def test_graphed_step_matches_eager_training():
    torch.manual_seed(0)
    raws = [make_batch("zinc", 32, seed=s) for s in (3, 4, 5)]
    cap = Capacity.covering(raws, slack=0.1)
    host = [pad_batch(r, cap, pin=True) for r in raws]
    m_eager = HL_HGCNN_zinc_dense_int3_pyr(**CTOR).to(DEV).train()
    m_graph = copy.deepcopy(m_eager)
    crit = torch.nn.L1Loss()

    def make_opt(m):
        return torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-3, fused=True, capturable=True)

    # graphed: 3 warm-up steps on host[0] happen inside the constructor
    bucket_g = FlatGradBucket(m_graph.parameters())
    stepper = GraphedTrainStep(m_graph, crit, make_opt(m_graph), bucket_g, host[0], DEV, warmup=3)
    assert stepper.launches_per_step > 50
    # eager reference: same 3 warm-up steps, unpadded batches
    bucket_e = FlatGradBucket(m_eager.parameters())
    opt_e = make_opt(m_eager)

    def eager_step(raw):
        b = batch_to(raw, DEV)
        bucket_e.zero()
        loss = crit(m_eager(b, device=DEV), b.y)
        loss.backward()
        opt_e.step()
        return loss.item()

    for _ in range(3):
        eager_step(raws[0])
    for i in (1, 2, 0, 1):
        stepper.batch.load(host[i])
        lg = float(stepper.step())
        le = eager_step(raws[i])
        assert abs(lg - le) < 1e-3 * max(1.0, abs(le)), (i, lg, le)
    for (n, a), (_, r) in zip(m_graph.named_parameters(), m_eager.named_parameters()):
        # biases in front of a BatchNorm have a zero true gradient; Adam turns their rounding noise
        # into +-lr steps, so only parameters with a real gradient are comparable
        if n.endswith(".weight"):
            assert float((a - r).norm()) < 5e-3 * float(r.norm()), n
