"""Two-lane issue (hlhgat_b200/lanes.py): node chain and edge chain on two CUDA streams.  The same kernels run in
the same per-lane order, so the forward pass (loss, BatchNorm running statistics) is BIT-IDENTICAL with the lanes
on or off.  In the backward pass autograd sums the gradient contributions that reach one tensor from two streams
in a different association than on one stream (measured: tools/lanes_debug.py, identical under
CUDA_LAUNCH_BLOCKING=1), so gradients agree to fp32 rounding (1e-5 of their norm here) -- and two-lane runs are
bit-identical to EACH OTHER, eager and inside the whole-step CUDA graph, where the two branches really overlap
(a race between the lanes would show up as run-to-run differences)."""
import copy
from types import SimpleNamespace

import pytest
import torch

import hlhgat_b200 as H
from hlhgat_b200 import lanes
from hlhgat_b200.lib import Hodge_ST_Model as M
from hlhgat_b200.parallel import FlatGradBucket
from hlhgat_b200.synthetic import make_batch, batch_to, make_tsp_batch
from hlhgat_b200.training import Capacity, pad_batch, pad_levels, GraphedTrainStep, StaticBatch
from hlhgat_b200.workloads import WORKLOADS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CTOR = dict(channels=[1, 2], filters=[32, 64], mlp_channels=[48], K=3, node_dim=21, edge_dim=3, keig=7)


@pytest.fixture(autouse=True)
def _lanes_off_afterwards():
    yield
    H.enable_lanes(False)


def _run(model, loss_of, use_lanes, reps=1):
    H.enable_lanes(use_lanes)
    out = []
    for _ in range(reps):
        m = copy.deepcopy(model)
        loss = loss_of(m)
        grads = torch.autograd.grad(loss, list(m.parameters()), allow_unused=True)
        torch.cuda.synchronize()
        out.append((loss.detach().clone(), [None if g is None else g.clone() for g in grads],
                    [b.clone() for b in m.buffers()]))
    H.enable_lanes(False)
    return out


def _assert_same(a, b, exact_grads=True):
    assert torch.equal(a[0], b[0]), (float(a[0]), float(b[0]))
    scale = max(float(y.abs().max()) for y in b[1] if y is not None)
    for x, y in zip(a[1], b[1]):
        assert (x is None) == (y is None)
        if x is not None and exact_grads:
            assert torch.equal(x, y)
        elif x is not None:
            # biases in front of a BatchNorm have a zero true gradient (pure rounding noise): floor on the scale
            # of the largest gradient entry of the model
            assert float((x - y).norm()) <= 1e-5 * float(y.norm()) + 1e-6 * scale * y.numel() ** 0.5
    for x, y in zip(a[2], b[2]):
        assert torch.equal(x, y)


def test_zinc_lanes_bit_identical_eager():
    torch.manual_seed(0)
    b = batch_to(make_batch("zinc", 96, seed=7), DEV)
    model = M.HL_HGCNN_zinc_dense_int3_pyr(**CTOR).to(DEV).train()

    def loss_of(m):
        return torch.nn.functional.l1_loss(m(b, device=DEV), b.y)

    off = _run(model, loss_of, False)[0]
    ons = _run(model, loss_of, True, reps=4)
    for on in ons:
        _assert_same(on, off, exact_grads=False)
        _assert_same(on, ons[0])
    assert lanes.active() is None


def _small(name):
    wl = copy.copy(WORKLOADS[name])
    ctor = dict(wl.ctor)
    ctor.update(channels=[1, 2, 1], filters=[32, 32, 64])
    if name != "tsp":
        ctor.update(mlp_channels=[48])
    wl.ctor = ctor
    return wl


@pytest.mark.parametrize("name,nb", [("peptides", 6), ("cifar", 4), ("tsp", 2)])
def test_other_models_lanes_bit_identical_eager(name, nb):
    torch.manual_seed(0)
    wl = _small(name)
    if name == "tsp":
        raw = make_tsp_batch(nb, seed=1, n=70, k=8)
        raw.y = raw.y.float()
    else:
        raw = wl.make(nb, 1)

    def dev_of(d):
        return SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in vars(d).items()})

    batch = [dev_of(l) for l in raw] if wl.levels > 1 else dev_of(raw)
    model = getattr(M, wl.model)(**wl.ctor).to(DEV).train()
    off = _run(model, lambda m: wl.loss(m, batch), False)[0]
    ons = _run(model, lambda m: wl.loss(m, batch), True, reps=3)
    for on in ons:
        _assert_same(on, off, exact_grads=False)
        _assert_same(on, ons[0])


def _check_graphed(results):
    """results = [off, on, on]: the two two-lane runs are bit-identical; training stays on the trajectory of the
    single-stream graph (forward bit-identity is checked by the eager tests: here the warm-up steps already
    moved the parameters by gradients that differ in the last bits)."""
    off, on1, on2 = results
    assert on1[0] == on2[0], (on1[0], on2[0])
    for a, b in zip(on1[1], on2[1]):
        assert torch.equal(a, b)
    for lo, ln in zip(off[0], on1[0]):
        assert abs(lo - ln) < 1e-3 * max(1.0, abs(lo)), (off[0], on1[0])
    for (n, _), a, b in zip(off[2], off[1], on1[1]):
        if n.endswith(".weight"):            # zero-gradient biases: Adam turns rounding noise into +-lr steps
            assert float((a - b).norm()) < 5e-3 * float(a.norm()), n


def test_graphed_step_lanes_bit_identical():
    """Whole-step CUDA graph with two parallel branches: deterministic, and on the trajectory of the single-stream
    graph after several optimizer steps on different batches."""
    torch.manual_seed(0)
    raws = [make_batch("zinc", 64, seed=s) for s in (3, 4, 5)]
    cap = Capacity.covering(raws, slack=0.1)
    host = [pad_batch(r, cap, pin=True) for r in raws]
    base = M.HL_HGCNN_zinc_dense_int3_pyr(**CTOR).to(DEV).train()
    crit = torch.nn.L1Loss()
    results = []
    for use in (False, True, True):
        H.enable_lanes(use)
        m = copy.deepcopy(base)
        bucket = FlatGradBucket(m.parameters())
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-3, fused=True, capturable=True)
        stepper = GraphedTrainStep(m, crit, opt, bucket, host[0], DEV, warmup=2)
        losses = []
        for i in (1, 2, 0, 1, 2):
            stepper.batch.load(host[i])
            losses.append(float(stepper.step()))
        torch.cuda.synchronize()
        results.append((losses, [p.detach().clone() for p in m.parameters()], list(m.named_parameters())))
        H.enable_lanes(False)
    _check_graphed(results)


def test_graphed_two_level_lanes_bit_identical():
    torch.manual_seed(0)
    wl = _small("peptides")
    raws = [wl.make(6, s) for s in (1, 2)]
    caps = [Capacity.covering([r[l] for r in raws], slack=0.1) for l in range(2)]
    host = [pad_levels(r, caps, pin=True, deg_eps=wl.deg_eps) for r in raws]
    base = getattr(M, wl.model)(**wl.ctor).to(DEV).train()
    results = []
    for use in (False, True, True):
        H.enable_lanes(use)
        m = copy.deepcopy(base)
        bucket = FlatGradBucket(m.parameters())
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)
        stepper = GraphedTrainStep(m, wl.loss, opt, bucket, host[0], DEV, warmup=2, loss_fn=True)
        losses = []
        for i in (1, 0, 1):
            stepper.batch.load(host[i])
            losses.append(float(stepper.step()))
        torch.cuda.synchronize()
        results.append((losses, [p.detach().clone() for p in m.parameters()], list(m.named_parameters())))
        H.enable_lanes(False)
    _check_graphed(results)


def test_weight_split_plan_replays_bit_identical():
    """functional.WeightSplitPlan: recorded weight splits re-issued in one launch give the same GEMM results, bit
    for bit, as splitting at the call site -- also after the weights were updated in place."""
    from hlhgat_b200 import functional as F_hl
    torch.manual_seed(0)
    a = torch.randn(1000, 96, device=DEV)
    a2 = torch.randn(1000, 40, device=DEV)          # 40 % 32 != 0: padded packing of [w1 | w2]
    g = torch.randn(1000, 64, device=DEV)
    w = torch.randn(64, 96, device=DEV) * 0.1
    w1 = torch.randn(64, 40, device=DEV) * 0.1
    w2 = torch.randn(64, 96, device=DEV) * 0.1

    def run():
        return (F_hl.dense(a, w), F_hl.dense(g, w, transpose_w=True), F_hl.dense2(a2, w1, a, w2),
                F_hl.dense2(a[:, :64], w[:, :64], a[:, 64:], w[:, 64:]))

    plan = F_hl.WeightSplitPlan()
    with plan:
        first = run()                                # records
    assert len(plan.entries) == 4 and plan._table_len == 6
    for t in (w, w1, w2):
        t.mul_(1.7).add_(0.01)
    ref = run()                                      # no plan: split at the call site
    with plan:
        got = run()                                  # replay: one hl_tf32_split_batch launch
        assert plan.ready is not None
    torch.cuda.synchronize()
    for x, y, z in zip(got, ref, first):
        assert torch.equal(x, y) and not torch.equal(x, z)


def test_project_then_transfer_equals_reference_order():
    """functional.enable_project_then_transfer: (1/D)|B1| (x_s W_a^T) + x_t W_b^T + b == [(1/D)|B1| x_s | x_t] W^T + b to
    fp32 rounding -- predictions, loss, every parameter gradient (weights of the split Linear included) and the
    BatchNorm running statistics; also under two-lane issue."""
    torch.manual_seed(0)
    b = batch_to(make_batch("zinc", 96, seed=9), DEV)
    ctor = dict(CTOR, channels=[2, 2])
    model = M.HL_HGCNN_zinc_dense_int3_pyr(**ctor).to(DEV).train()

    def loss_of(m):
        return torch.nn.functional.l1_loss(m(b, device=DEV), b.y)

    ref = _run(model, loss_of, False)[0]
    for use_lanes in (False, True):
        H.enable_project_then_transfer(True)
        try:
            got = _run(model, loss_of, use_lanes)[0]
        finally:
            H.enable_project_then_transfer(False)
        assert abs(float(got[0]) - float(ref[0])) < 1e-5 * max(1.0, abs(float(ref[0])))
        scale = max(float(y.abs().max()) for y in ref[1] if y is not None)
        for (n, _), x, y in zip(model.named_parameters(), got[1], ref[1]):
            assert (x is None) == (y is None), n
            if x is not None:
                assert float((x - y).norm()) <= 1e-4 * float(y.norm()) + 1e-6 * scale * y.numel() ** 0.5, n
        for x, y in zip(got[2], ref[2]):
            assert torch.allclose(x.float(), y.float(), rtol=1e-4, atol=1e-6)
