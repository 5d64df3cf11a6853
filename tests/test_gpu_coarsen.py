"""GPU multi-level graph coarsening (hlhgat_b200.coarsen.mlgc_batch) against (i) the golden vectors of the
reference's MLGC run with the deterministic graclus stand-in (tests/golden/pool.pt), (ii) the host restatement
used for the synthetic two-level batches, on whole mini-batches of config-shaped graphs, and (iii) the
heavy-edge (weighted) variant of MLGC_weighted."""
import sys
import os

import numpy as np
import pytest
import torch

import hlhgat_b200  # noqa: F401
from hlhgat_b200.coarsen import mlgc_batch
from hlhgat_b200.construct import build_simplex_batch
from hlhgat_b200.synthetic import SHAPES, _knn_graph, _tree_plus_chords, coarsen, simplex_graph
from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fine(eis, ns):
    off, src, dst = 0, [], []
    for ei, n in zip(eis, ns):
        src += [ei[0] + off, ei[1] + off]
        dst += [ei[1] + off, ei[0] + off]
        off += n
    return build_simplex_batch(torch.cat(src).to(DEV), torch.cat(dst).to(DEV), torch.tensor(ns))


def test_mlgc_vs_golden_reference():
    c = load_golden("pool.pt")
    g = c["fine"]
    sb = build_simplex_batch(g["ei_dir"][0].to(DEV), g["ei_dir"][1].to(DEV), torch.tensor([g["n"]]))
    coarse, c_node, c_edge = mlgc_batch(sb)
    assert torch.equal(c_node.cpu().long(), c["c_node"].long())
    assert torch.equal(c_edge.cpu(), c["c_edge"])
    ref = c["coarse"]
    assert torch.equal(coarse.edge_index.cpu(), ref["edge_index"])
    for side in ("t", "s"):
        ei, ew = coarse.coo(side)
        assert torch.equal(ei.cpu(), ref[f"edge_index_{side}"]), side
        assert torch.allclose(ew.cpu(), ref[f"edge_weight_{side}"], rtol=3e-7, atol=0), side
    assert int(coarse.num_node1[0]) == int(ref["n"]) and int(coarse.num_edge1[0]) == int(ref["e"])


@pytest.mark.parametrize("shape,batch", [("zinc", 48), ("peptides", 8), ("cifar", 6), ("tsp", 2)])
def test_mlgc_batch_vs_host_restatement(shape, batch):
    rng = np.random.default_rng(11)
    n_lo, n_hi, kind, param, _, _ = SHAPES[shape]
    eis, ns, refs = [], [], []
    for _ in range(batch):
        n = int(rng.integers(n_lo, n_hi + 1))
        ei = _tree_plus_chords(rng, n, param) if kind == "tree" else _knn_graph(rng, n, param)
        eis.append(torch.from_numpy(np.ascontiguousarray(ei)))
        ns.append(n)
        refs.append(coarsen(ei, n))
    coarse, c_node, c_edge = mlgc_batch(_fine(eis, ns))
    assert torch.equal(c_node.view(-1).cpu().long(), torch.from_numpy(np.concatenate([r[2] for r in refs])))
    assert torch.equal(c_edge.view(-1).cpu(), torch.from_numpy(np.concatenate([r[3] for r in refs])))
    assert coarse.num_node1.cpu().tolist() == [r[1] for r in refs]
    assert coarse.num_edge1.cpu().tolist() == [r[0].shape[1] for r in refs]
    off, cols = 0, []
    for r in refs:
        cols.append(torch.from_numpy(r[0]) + off)
        off += r[1]
    assert torch.equal(coarse.edge_index.cpu(), torch.cat(cols, 1))
    # coarse operators of the first and the last graph against the host construction formulas
    it, wt = coarse.coo("t")
    is_, ws = coarse.coo("s")
    g0 = simplex_graph(refs[0][0], refs[0][1])
    nt, ns_ = g0["edge_index_t"].shape[1], g0["edge_index_s"].shape[1]
    assert torch.equal(it[:, :nt].cpu(), torch.from_numpy(g0["edge_index_t"]))
    assert torch.equal(is_[:, :ns_].cpu(), torch.from_numpy(g0["edge_index_s"]))
    assert torch.allclose(wt[:nt].cpu(), torch.from_numpy(g0["edge_weight_t"]), rtol=2e-6, atol=0)
    assert torch.allclose(ws[:ns_].cpu(), torch.from_numpy(g0["edge_weight_s"]), rtol=2e-6, atol=0)


def test_weighted_matching_vs_graclus_stand_in():
    """MLGC_weighted (lib/Hodge_Dataset.py:309-311): heavy-edge matching on edge weights exp(-x^2)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pyg_shim"))
    from torch_cluster import graclus_cluster
    from hlhgat_b200.coarsen import greedy_matching
    rng = np.random.default_rng(5)
    eis, ns, ws, want = [], [], [], []
    off = 0
    for _ in range(12):
        n = int(rng.integers(20, 60))
        ei = torch.from_numpy(np.ascontiguousarray(_knn_graph(rng, n, 4)))
        w = torch.exp(-torch.from_numpy(rng.random(ei.shape[1]).astype(np.float32)) ** 2)
        w[rng.integers(0, ei.shape[1], 5)] = 0.75                        # exact ties
        row, col = torch.cat([ei[0], ei[1]]), torch.cat([ei[1], ei[0]])
        order = torch.argsort(row * n + col)                              # to_undirected: coalesced order
        want.append(graclus_cluster(row[order], col[order], torch.cat([w, w])[order], n) + off)
        eis.append(ei)
        ns.append(n)
        ws.append(w)
        off += n
    got = greedy_matching(_fine(eis, ns), torch.cat(ws))
    assert torch.equal(got.cpu().long(), torch.cat(want))


def test_attpool_model_on_gpu_built_two_level_batch_equals_host_built():
    """raw graphs -> GPU construction + GPU MLGC -> model == the same model on the host-built two-level batch."""
    from types import SimpleNamespace
    from hlhgat_b200.lib import Hodge_ST_Model as M
    from hlhgat_b200.lib.Hodge_Dataset import two_level_batch_from_graphs
    from hlhgat_b200.synthetic import make_multilevel_batch
    host = make_multilevel_batch("cifar", 5, seed=4, node_dim=8, edge_dim=6)
    lv0 = host[0]
    ns = lv0.num_node1.tolist()
    es = lv0.num_edge1.tolist()
    eis, n_off, e_off = [], 0, 0
    for n, e in zip(ns, es):
        eis.append(lv0.edge_index[:, e_off:e_off + e] - n_off)
        n_off, e_off = n_off + n, e_off + e
    datas = two_level_batch_from_graphs(eis, ns, lv0.x_t[:, 1:], lv0.x_s[:, 1:], device=DEV)
    assert torch.equal(datas[0].x_t[:, 0].cpu(), lv0.x_t[:, 0]) and torch.equal(datas[0].x_s[:, 0].cpu(), lv0.x_s[:, 0])
    assert torch.equal(datas[1].edge_index.cpu(), host[1].edge_index)
    torch.manual_seed(0)
    ctor = dict(channels=[1, 1, 1], filters=[32, 32, 64], mlp_channels=[32], K=3, node_dim=4, edge_dim=2, keig=4, pool_loc=1, num_classes=3)
    model = M.HL_HGCNN_pepfunc_dense_int3_attpool(**ctor).to(DEV).train()
    dev_host = [SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in vars(d).items()}) for d in host]
    a = model(datas, device=DEV)
    b = model(dev_host, device=DEV)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), float((a - b).abs().max())
