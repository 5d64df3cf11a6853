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
    bucket.all_reduce_mean()
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
