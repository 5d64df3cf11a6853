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
        # four Adam steps apart: the padded batch splits the weight-gradient row range differently (other rounding), and
        # Adam turns the rounding noise of zero-gradient biases into +-lr steps -- the trajectories agree to ~1e-3
        assert abs(lg - le) < 3e-3 * max(1.0, abs(le)), (i, lg, le)
    for (n, a), (_, r) in zip(m_graph.named_parameters(), m_eager.named_parameters()):
        # biases in front of a BatchNorm have a zero true gradient; Adam turns their rounding noise
        # into +-lr steps, so only parameters with a real gradient are comparable
        if n.endswith(".weight"):
            assert float((a - r).norm()) < 5e-3 * float(r.norm()), n


# ---------------------------------------------------------------------------------------------
# the other workloads (two-level attpool batches, per-edge TSP targets) through the same machinery
# ---------------------------------------------------------------------------------------------
def _small(wl_name):
    from hlhgat_b200.workloads import WORKLOADS
    import copy as _c
    wl = _c.copy(WORKLOADS[wl_name])
    ctor = dict(wl.ctor)
    ctor.update(channels=[1, 1, 1], filters=[32, 32, 64])
    if wl_name != "tsp":
        ctor.update(mlp_channels=[48])
    wl.ctor = ctor
    return wl


@pytest.mark.parametrize("name,nb", [("peptides", 6), ("cifar", 4), ("tsp", 2)])
def test_padded_equals_unpadded_and_graph_equals_eager_other_workloads(name, nb):
    from hlhgat_b200.lib import Hodge_ST_Model as M
    from hlhgat_b200.training import pad_levels
    from types import SimpleNamespace
    torch.manual_seed(0)
    wl = _small(name)
    if name == "tsp":
        from hlhgat_b200.synthetic import make_tsp_batch
        raws = []
        for s in (1, 2):
            b = make_tsp_batch(nb, seed=s, n=60 + 10 * s, k=8)
            b.y = b.y.float()
            raws.append(b)
    else:
        raws = [wl.make(nb, s) for s in (1, 2)]

    def dev_of(b):
        return SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in vars(b).items()})

    if wl.levels > 1:
        caps = [Capacity.covering([r[l] for r in raws], slack=0.1) for l in range(2)]
        host = [pad_levels(r, caps, pin=True, deg_eps=wl.deg_eps) for r in raws]
        plain = [[dev_of(l) for l in r] for r in raws]
    else:
        caps = [Capacity.covering(raws, slack=0.1)]
        host = [pad_batch(r, caps[0], pin=True, deg_eps=wl.deg_eps) for r in raws]
        plain = [dev_of(r) for r in raws]
    model = getattr(M, wl.model)(**wl.ctor).to(DEV).train()
    ref = copy.deepcopy(model)
    # (1) padded == unpadded, eager
    loss_ref = wl.loss(ref, plain[0])
    g_ref = torch.autograd.grad(loss_ref, list(ref.parameters()), allow_unused=True)
    sb = StaticBatch(host[0], DEV)
    loss_p = wl.loss(model, sb)
    assert abs(float(loss_p) - float(loss_ref)) < 1e-4 * max(1.0, abs(float(loss_ref))), (float(loss_p), float(loss_ref))
    g = torch.autograd.grad(loss_p, list(model.parameters()), allow_unused=True)
    for (n, _), a, r in zip(model.named_parameters(), g, g_ref):
        assert (a is None) == (r is None), n
        if a is not None:
            assert float((a - r).norm()) < 2e-3 * float(r.norm()) + 1e-6 * r.numel() ** 0.5 * max(1.0, float(loss_ref)), n
    # (2) graph replay == eager on the padded batch
    m_graph = copy.deepcopy(ref)
    bucket = FlatGradBucket(m_graph.parameters())
    opt = torch.optim.SGD(m_graph.parameters(), lr=0.0)
    stepper = GraphedTrainStep(m_graph, wl.loss, opt, bucket, host[0], DEV, warmup=2, loss_fn=True)
    stepper.batch.load(host[1])
    lg = float(stepper.step())
    le = float(wl.loss(copy.deepcopy(ref), StaticBatch(host[1], DEV)))
    assert abs(lg - le) < 1e-4 * max(1.0, abs(le)), (lg, le)


def test_batch_prefetcher_delivers_every_batch_in_order():
    """training.BatchPrefetcher: H2D on a copy stream into two staging slots, device-to-device into the step's static
    buffers; the consumer must see batch i in step i whatever the overlap."""
    from hlhgat_b200.training import BatchPrefetcher
    raws = [make_batch("zinc", 16, seed=s) for s in range(5)]
    cap = Capacity.covering(raws, slack=0.1)
    host = [pad_batch(r, cap, pin=True) for r in raws]
    target = StaticBatch(host[0], DEV)
    pre = BatchPrefetcher(target, host[0], DEV)
    pre.prefetch(host[0], 0)
    burn = torch.randn(2048, 2048, device=DEV)
    for i in range(12):
        pre.swap_in(i % 2)
        pre.prefetch(host[(i + 1) % 5], (i + 1) % 2)
        got = target.x_t.clone(), target.edge_index_s.clone(), target.n_valid_nodes.clone()
        burn = burn @ burn * 1e-3                               # keep the compute stream busy while the next copy runs
        want = host[i % 5]
        assert torch.equal(got[0].cpu(), want.x_t) and torch.equal(got[1].cpu(), want.edge_index_s)
        assert int(got[2]) == int(want.n_valid_nodes)


def test_flat_adam_matches_torch_adam():
    """parallel.FlatAdam (hl_adam_flat: one streaming kernel over flat parameter / gradient / moment buffers) against
    torch.optim.Adam with the training scripts' settings (lr 1e-3, weight decay 1e-3) on identical gradients, ten steps."""
    from hlhgat_b200.parallel import FlatAdam
    torch.manual_seed(0)
    shapes = [(64, 28), (64,), (7,), (128, 130), (1, 3), (33,)]                 # odd sizes: the unaligned tail path too
    ref_p = [torch.nn.Parameter(torch.randn(*s, device=DEV)) for s in shapes]
    my_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=1e-3)
    bucket = FlatGradBucket(my_p)
    opt = FlatAdam(bucket, lr=1e-3, weight_decay=1e-3)
    assert all(p.data_ptr() >= opt.flat_params.data_ptr() for p in my_p)        # parameters now live in the flat buffer
    for step in range(10):
        grads = [torch.randn_like(p) * (1.0 + step) for p in ref_p]
        for p, q, g in zip(ref_p, my_p, grads):
            p.grad = g.clone()
            q.grad.copy_(g * 4.0)                                               # as if summed over 4 ranks ...
        ref.step()
        opt.step(grad_scale=0.25)                                               # ... and averaged inside the kernel
    for p, q in zip(ref_p, my_p):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), float((p - q).abs().max())
    assert float(opt.state[0]) == 10.0
