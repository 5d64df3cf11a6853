"""Host-side logic of the N>1 path on CPU: world_size-2 gloo run of the flat gradient bucket
(hl-hgat_b200/parallel.py) against the single-process average, and the graph sharding rule."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTOR = dict(channels=[1, 1], filters=[8, 12], mlp_channels=[], K=2, node_dim=4, edge_dim=2, keig=3)


def _grads_for(seed_batch, state):
    sys.path.insert(0, ROOT)
    from oracle import hodge_oracle as O
    from hlhgat_b200.synthetic import make_batch
    model = O.HL_HGCNN_zinc_dense_int3_pyr(**CTOR).train()
    model.load_state_dict(state)
    b = make_batch("zinc", 6, seed=seed_batch, node_dim=7, edge_dim=5)
    return model, b


def _worker(rank, world, init_file, state, out_file):
    sys.path.insert(0, ROOT)
    import hlhgat_b200  # noqa: F401
    from hlhgat_b200.parallel import FlatGradBucket, broadcast_parameters
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.manual_seed(rank)                              # different init per rank: broadcast must fix it
    model, b = _grads_for(10 + rank, state)
    if rank == 1:
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    broadcast_parameters(model, src=0)
    bucket = FlatGradBucket(model.parameters())
    bucket.zero()
    torch.nn.functional.l1_loss(model(b), b.y).backward()
    summed = bucket.flat.clone()
    bucket.all_reduce_mean()
    # the SUM-only variant leaves the 1 / world_size to the optimizer (parallel.FlatAdam reads `pending_scale`)
    bucket.flat.copy_(summed)
    bucket.all_reduce_sum()
    assert bucket.pending_scale == 1.0 / world
    bucket.flat.mul_(bucket.pending_scale)
    if rank == 0:
        torch.save(bucket.flat.clone(), out_file)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2_gloo():
    sys.path.insert(0, ROOT)
    from oracle import hodge_oracle as O
    torch.manual_seed(0)
    state = O.HL_HGCNN_zinc_dense_int3_pyr(**CTOR).state_dict()
    with tempfile.TemporaryDirectory() as d:
        init_file, out_file = os.path.join(d, "init"), os.path.join(d, "flat.pt")
        mp.spawn(_worker, args=(2, init_file, state, out_file), nprocs=2, join=True)
        got = torch.load(out_file)
    flats = []
    for r in range(2):
        model, b = _grads_for(10 + r, state)
        torch.nn.functional.l1_loss(model(b), b.y).backward()
        flats.append(torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad]))
    want = (flats[0] + flats[1]) / 2
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-7)


def test_shard_graphs_balances_cost():
    sys.path.insert(0, ROOT)
    from hlhgat_b200.parallel import shard_graphs
    costs = [100, 1, 1, 1, 50, 50, 3, 97]
    parts = shard_graphs(costs, 2)
    assert sorted(i for p in parts for i in p) == list(range(8))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 3
    assert shard_graphs([5, 4], 4) == [[0], [1], [], []]


# ---------------------------------------------------------------------------------------------
# BatchNorm statistics over all ranks (parallel.combine_bn_stats / reduce_bn_sums): the exchange and the merge
# formulas, world size 2 on gloo, against torch's batch_norm + autograd on the concatenated batch.  The per-rank
# phases (what hl_bn_stats / hl_bn_bwd_sums / hl_bn_bwd_apply compute on the GPU) are restated in torch here.
# ---------------------------------------------------------------------------------------------
def _syncbn_worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    import hlhgat_b200  # noqa: F401
    from hlhgat_b200 import parallel as P
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    x_all = torch.randn(37, 6, generator=g) * 3 + 100.0          # large mean: E[x^2] - mean^2 would cancel in fp32
    dy_all = torch.randn(37, 6, generator=g)
    gamma = torch.rand(6, generator=g) + 0.5
    rows = slice(0, 23) if rank == 0 else slice(23, 37)          # unequal shards
    x, dy = x_all[rows], dy_all[rows]
    eps, slope = 1e-5, 0.1
    local = torch.cat([x.mean(0), x.var(0, unbiased=False)])
    stats, total = P.combine_bn_stats(local, torch.tensor([float(x.shape[0])]))
    mean, rstd = stats[:6], (stats[6:] + eps).rsqrt()
    xhat = (x - mean) * rstd
    z = xhat * gamma
    y = torch.where(z > 0, z, z * slope)
    dz = torch.where(y > 0, dy, dy * slope)
    sums = torch.cat([dz.sum(0), (dz * xhat).sum(0)])
    gs = P.reduce_bn_sums(sums)
    dx = gamma * rstd * (dz - gs[:6] / total - xhat * gs[6:] / total)
    torch.save({"stats": stats, "total": total, "y": y, "dx": dx, "dgamma": sums[6:], "dbeta": sums[:6]},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sync_batchnorm_exchange_world2_gloo():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_syncbn_worker, args=(2, os.path.join(d, "init"), d), nprocs=2, join=True)
        r = [torch.load(os.path.join(d, f"r{k}.pt")) for k in range(2)]
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(37, 6, generator=g) * 3 + 100.0).requires_grad_(True)
    dy = torch.randn(37, 6, generator=g)
    gamma = (torch.rand(6, generator=g) + 0.5).requires_grad_(True)
    beta = torch.zeros(6, requires_grad=True)
    y = torch.nn.functional.leaky_relu(torch.nn.functional.batch_norm(x, None, None, gamma, beta, True, 0.0, 1e-5), 0.1)
    dx, dgamma, dbeta = torch.autograd.grad(y, (x, gamma, beta), dy)
    assert float(r[0]["total"]) == 37.0
    assert torch.equal(r[0]["stats"], r[1]["stats"])
    assert torch.allclose(r[0]["stats"][:6], x.detach().mean(0), rtol=1e-6)
    assert torch.allclose(r[0]["stats"][6:], x.detach().var(0, unbiased=False), rtol=1e-4)
    assert torch.allclose(torch.cat([r[0]["y"], r[1]["y"]]), y.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(torch.cat([r[0]["dx"], r[1]["dx"]]), dx, rtol=1e-3, atol=1e-5)
    assert torch.allclose(r[0]["dgamma"] + r[1]["dgamma"], dgamma, rtol=1e-3, atol=1e-5)   # summed by the gradient all-reduce
    assert torch.allclose(r[0]["dbeta"] + r[1]["dbeta"], dbeta, rtol=1e-4, atol=1e-5)
