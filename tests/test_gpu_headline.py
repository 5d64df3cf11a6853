"""Parity AT the headline configuration (BASELINE.json configs[1]): the path bench.py times -- 1024 ZINC-shaped graphs
padded to a fixed capacity, node lane / edge lane on two streams, forward + backward replayed as one CUDA graph with the
weight gradients accumulated straight into the flat bucket -- against the CPU oracle on the same batch.

Forward: every prediction within rtol 1e-4 of the fp32 oracle AND of the fp64 oracle (north-star bar).
Gradients: the bar is PINNED, not narrated: the fp32 CPU oracle's own distance from its fp64 run is measured on the same
batch, and the CUDA path has to land within a small factor of it (both are fp32 evaluations of the same function; through
38 ReLU(BatchNorm) layers their masks flip for elements within rounding distance of zero)."""
import copy

import pytest
import torch

import hlhgat_b200 as H
from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
from hlhgat_b200.parallel import FlatGradBucket
from hlhgat_b200.synthetic import make_batch
from hlhgat_b200.training import Capacity, pad_batch, GraphedTrainStep
from hlhgat_b200.workloads import WORKLOADS
from oracle import hodge_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a.double() - b.double()).norm()) / max(float(b.double().norm()), 1e-30)


def test_headline_1024_graph_replayed_two_lane_step_vs_oracle():
    wl = WORKLOADS["zinc"]
    torch.manual_seed(0)
    ref = O.HL_HGCNN_zinc_dense_int3_pyr(**wl.ctor).train()
    raws = [make_batch("zinc", wl.batch, seed=s) for s in (7, 8)]
    b = raws[1]
    # CPU oracle, fp32 and fp64, on the unpadded batch
    ref64 = copy.deepcopy(ref).double()
    pred32 = ref(b)
    g32 = torch.autograd.grad(torch.nn.functional.l1_loss(pred32, b.y), list(ref.parameters()), allow_unused=True)
    b64 = copy.copy(b)
    for k in ("x_t", "x_s", "y", "edge_weight_t", "edge_weight_s"):
        setattr(b64, k, getattr(b, k).double())
    pred64 = ref64(b64)
    g64 = torch.autograd.grad(torch.nn.functional.l1_loss(pred64, b64.y), list(ref64.parameters()), allow_unused=True)

    # the bench path
    model = HL_HGCNN_zinc_dense_int3_pyr(**wl.ctor).to(DEV).train()
    model.load_state_dict(ref.state_dict(), strict=True)
    cap = Capacity.covering(raws)
    host = [pad_batch(r, cap, pin=True, deg_eps=wl.deg_eps) for r in raws]
    stash = torch.zeros(wl.batch, 1, device=DEV)

    def loss_fn(m, batch):
        out = m(batch, device=DEV)
        stash.copy_(out[: batch.num_graphs].detach())
        return torch.nn.functional.l1_loss(out[: batch.num_graphs], batch.y)

    H.enable_lanes(True)
    try:
        bucket = FlatGradBucket(model.parameters())
        opt = torch.optim.SGD(model.parameters(), lr=0.0)
        stepper = GraphedTrainStep(model, loss_fn, opt, bucket, host[0], DEV, warmup=3, loss_fn=True)
        stepper.batch.load(host[1])
        loss = float(stepper.step())
        torch.cuda.synchronize()
    finally:
        H.enable_lanes(False)
    pred = stash.cpu()
    assert torch.allclose(pred, pred32.detach(), rtol=1e-4, atol=1e-4), (pred - pred32).abs().max()
    assert torch.allclose(pred, pred64.detach().float(), rtol=1e-4, atol=1e-4), (pred - pred64.float()).abs().max()
    assert abs(loss - float(torch.nn.functional.l1_loss(pred64, b64.y))) < 1e-4

    d_ours, d_cpu = [], []
    gmax = max(float(r.norm()) for r in g64 if r is not None)
    for (n, p), a32, r64 in zip(model.named_parameters(), g32, g64):
        if r64 is None:
            continue
        if float(r64.norm()) < 1e-6 * gmax:          # biases feeding a BatchNorm: exactly-zero true gradient
            assert float(p.grad.norm()) < 1e-5 * max(1.0, gmax), n
            continue
        d_ours.append(_rel(p.grad.cpu(), r64))
        d_cpu.append(_rel(a32, r64))
    mean_o, mean_c = sum(d_ours) / len(d_ours), sum(d_cpu) / len(d_cpu)
    print(f"gradient distance from the fp64 oracle over {len(d_ours)} tensors: CUDA path mean {mean_o:.2e} max {max(d_ours):.2e}; "
          f"fp32 CPU oracle mean {mean_c:.2e} max {max(d_cpu):.2e}")
    # the fp32 CPU oracle itself is this far from fp64; the CUDA path must be of the same scale
    assert mean_o < 3.0 * mean_c + 1e-4, (mean_o, mean_c)
    assert max(d_ours) < 3.0 * max(d_cpu) + 1e-3, (max(d_ours), max(d_cpu))
    assert mean_o < 1e-2 and max(d_ours) < 3e-2
