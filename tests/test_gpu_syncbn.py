"""BatchNorm statistics over all data-parallel ranks (parallel.enable_sync_batchnorm; SURVEY.md section 8e, caveat 1):
two ranks (NCCL, one GPU each), each with half of the graphs, reproduce the single-GPU run of the GLOBAL batch --
predictions, averaged gradients and running statistics -- which per-rank statistics (plain DP) do not.
Needs 2 GPUs (skipped on a 1-GPU box)."""
import copy
import os
import sys
import tempfile
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]
CTOR = dict(channels=[1, 1], filters=[32, 64], mlp_channels=[48], K=2, node_dim=21, edge_dim=3, keig=7)


def _merge(a, b):
    """block-diagonal union of two collated batches (PairData.__inc__ offsets, lib/Hodge_Dataset.py:40-48)"""
    n, e = a.x_t.shape[0], a.x_s.shape[0]
    m = SimpleNamespace(num_graphs=a.num_graphs + b.num_graphs)
    for k in ("x_t", "x_s", "y", "edge_weight_t", "edge_weight_s", "num_node1", "num_edge1"):
        setattr(m, k, torch.cat([getattr(a, k), getattr(b, k)]))
    m.edge_index = torch.cat([a.edge_index, b.edge_index + n], 1)
    m.edge_index_t = torch.cat([a.edge_index_t, b.edge_index_t + n], 1)
    m.edge_index_s = torch.cat([a.edge_index_s, b.edge_index_s + e], 1)
    return m


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    import hlhgat_b200 as H
    from hlhgat_b200 import parallel as P
    from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
    from hlhgat_b200.synthetic import make_batch, batch_to
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world,
                            device_id=torch.device(dev))
    torch.manual_seed(0)
    base = HL_HGCNN_zinc_dense_int3_pyr(**CTOR).to(dev).train()
    P.broadcast_parameters(base)
    raws = [make_batch("zinc", 24, seed=11), make_batch("zinc", 24, seed=12)]
    mine = batch_to(raws[rank], dev)
    res = {}
    for name, sync, use_lanes in (("dp", False, False), ("sync", True, False), ("sync_lanes", True, True)):
        P.enable_sync_batchnorm(sync)
        H.enable_lanes(use_lanes)
        m = copy.deepcopy(base)
        bucket = P.FlatGradBucket(m.parameters())
        bucket.zero()
        pred = m(mine, device=dev)
        torch.nn.functional.l1_loss(pred, mine.y).backward()
        bucket.all_reduce_mean()
        torch.cuda.synchronize()
        res[name] = {"pred": pred.detach().cpu(), "flat": bucket.flat.detach().cpu().clone(),
                     "buffers": {k: v.detach().cpu().clone() for k, v in m.named_buffers()}}
    P.enable_sync_batchnorm(False)
    H.enable_lanes(False)
    if rank == 0:                                  # the single-GPU run of the global batch
        m = copy.deepcopy(base)
        bucket = P.FlatGradBucket(m.parameters())
        bucket.zero()
        g = batch_to(_merge(*raws), dev)
        pred = m(g, device=dev)
        torch.nn.functional.l1_loss(pred, g.y).backward()
        torch.cuda.synchronize()
        res["global"] = {"pred": pred.detach().cpu(), "flat": bucket.flat.detach().cpu().clone(),
                         "buffers": {k: v.detach().cpu().clone() for k, v in m.named_buffers()}}
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sync_batchnorm_two_ranks_equal_single_gpu_global_batch():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "init"), d), nprocs=2, join=True)
        r = [torch.load(os.path.join(d, f"r{k}.pt")) for k in range(2)]
    glob = r[0]["global"]
    for name in ("sync", "sync_lanes"):
        pred = torch.cat([r[0][name]["pred"], r[1][name]["pred"]])
        assert torch.allclose(pred, glob["pred"], rtol=1e-4, atol=1e-4 * float(glob["pred"].abs().max())), name
        a, b = r[0][name]["flat"], glob["flat"]
        assert torch.equal(a, r[1][name]["flat"])
        assert float((a - b).norm()) < 1e-3 * float(b.norm()), (name, float((a - b).norm()), float(b.norm()))
        for k, v in glob["buffers"].items():
            assert torch.allclose(r[0][name]["buffers"][k].float(), v.float(), rtol=1e-4, atol=1e-6), (name, k)
    # plain data parallelism (per-rank statistics) is a different function of the global batch
    pred_dp = torch.cat([r[0]["dp"]["pred"], r[1]["dp"]["pred"]])
    assert not torch.allclose(pred_dp, glob["pred"], rtol=1e-4, atol=1e-4 * float(glob["pred"].abs().max()))
