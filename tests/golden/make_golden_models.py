"""Generate tests/golden/models.pt by running the UNMODIFIED reference model classes of the other
BASELINE.json configs (TSP pyr, CIFAR10-superpixel attpool, peptides-func attpool) forward + backward.

Runs only in the build container (needs /root/reference, read-only).  The third-party primitives are the
pure-torch stand-ins of `oracle/pyg_shim`; the control flow is the reference's own:
  * lib/Hodge_ST_Model.py:756-852   HL_HGCNN_TSP_dense_int3_pyr
  * lib/Hodge_ST_Model.py:958-1091  HL_HGCNN_CIFAR10SP_dense_int3_attpool
  * main_pepfunc_HL_HGCNN_dense_int3_attpool.py:36-168  HL_HGCNN_pepfunc_dense_int3_attpool (the class is
    defined inline in a script that parses argv at import, so its source lines are exec'd verbatim)
  * lib/Hodge_Dataset.py:241-295    MLGC (level-1 graphs; graclus shimmed deterministically)

    python tests/golden/make_golden_models.py      # rewrites tests/golden/models.pt (deterministic)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

from make_golden import RC, RD, RM, Batch, rand_graph, ref_construct  # noqa: E402  (sets up the shim paths)

KEYS = ("x_t", "x_s", "edge_index", "edge_index_t", "edge_index_s", "edge_weight_t", "edge_weight_s",
        "num_node1", "num_edge1")


def load_pepfunc_class():
    """exec the class statement of the reference script verbatim (lines 36-168) in the namespace the
    script itself builds with its imports."""
    path = "/root/reference/main_pepfunc_HL_HGCNN_dense_int3_attpool.py"
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("class HL_HGCNN_pepfunc_dense_int3_attpool"))
    end = next(i for i in range(start + 1, len(lines)) if lines[i] and not lines[i][0].isspace() and not lines[i].startswith("#"))
    src = "\n".join(lines[start:end])
    import torch.nn as nn
    import torch_geometric.nn as gnn
    from torch_geometric.nn import global_mean_pool
    from torch_geometric.utils import degree
    from torch_scatter import scatter_mean
    from torch.nn import Linear, Dropout
    ns = dict(torch=torch, nn=nn, gnn=gnn, global_mean_pool=global_mean_pool, degree=degree,
              scatter_mean=scatter_mean, Linear=Linear, Dropout=Dropout, HodgeLaguerreConv=RC.HodgeLaguerreConv,
              NodeEdgeInt=RC.NodeEdgeInt, adj2par1=RD.adj2par1)
    exec(compile(src, path, "exec"), ns)
    return ns["HL_HGCNN_pepfunc_dense_int3_attpool"]


def pair(g, x_t, x_s, y=None):
    n, e = g["n"], g["edge_index"].shape[1]
    d = RD.PairData(x_s=x_s, edge_index_s=g["edge_index_s"], edge_weight_s=g["edge_weight_s"],
                    x_t=x_t, edge_index_t=g["edge_index_t"], edge_weight_t=g["edge_weight_t"], y=y)
    d.num_node1, d.num_edge1, d.num_nodes, d.edge_index = n, e, n, g["edge_index"]
    return d


def two_level_batches(rng, n_graphs, n_lo, n_hi, extra, fn, fe, n_cls):
    """per sample [fine, coarse] exactly like the datasets' get() (lib/Hodge_Dataset.py:866-870): coarse =
    MLGC(fine), cluster ids prepended as column 0 of x_t / x_s; then the DataLoader's list collation."""
    fines, coarses = [], []
    for _ in range(n_graphs):
        n = int(rng.integers(n_lo, n_hi))
        g = ref_construct(rand_graph(rng, n, extra), n)
        e = g["edge_index"].shape[1]
        fine = pair(g, torch.randn(n, fn), torch.randn(e, fe), y=torch.randint(0, n_cls, (1,)))
        coarse, c_node, c_edge = RD.MLGC(fine)
        fine.x_t = torch.cat([c_node.float(), fine.x_t], -1)
        fine.x_s = torch.cat([c_edge, fine.x_s], -1)
        fines.append(fine)
        coarses.append(coarse)
    return [Batch.from_data_list(fines), Batch.from_data_list(coarses)]


def run(model, call, target_loss):
    model.train()
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    pred = call(model)
    loss = target_loss(pred)
    gr = torch.autograd.grad(loss, [p for p in model.parameters()], allow_unused=True)
    return dict(state=state0, pred=pred.detach(), loss=loss.detach(),
                grads={n_: t for (n_, _), t in zip(model.named_parameters(), gr)})


def tsp_case(rng):
    torch.manual_seed(21)
    graphs = []
    for _ in range(3):
        n = int(rng.integers(10, 15))
        g = ref_construct(rand_graph(rng, n, 8), n)
        e = g["edge_index"].shape[1]
        x_s = torch.cat([torch.rand(e, 1), (torch.rand(e, 1) > 0.2).float()], -1)     # [length, edge_mask]
        graphs.append(pair(g, torch.rand(n, 2), x_s))
    batch = Batch.from_data_list(graphs)
    ctor = dict(channels=[1, 2], filters=[8, 16], mlp_channels=[12], K=3, node_dim=2, edge_dim=1, num_classes=2)
    torch.manual_seed(22)
    model = RM.HL_HGCNN_TSP_dense_int3_pyr(**ctor)
    w = torch.randn(batch.x_s.shape[0], 2)
    res = run(model, lambda m: m(batch, device="cpu")[0], lambda p: (p * w).sum() / p.shape[0])
    return dict(batch={k: batch[k] for k in KEYS}, ctor=ctor, w=w, **res)


def attpool_case(rng, cls, ctor, seed):
    torch.manual_seed(seed)
    datas = two_level_batches(rng, 4, 12, 18, 5, ctor["node_dim"] + ctor["keig"], ctor["edge_dim"] + ctor["keig"], 3)
    saved = [{k: b[k].clone() for k in KEYS} for b in datas]
    torch.manual_seed(seed + 1)
    model = cls(**ctor)
    w = torch.randn(4, ctor["num_classes"])
    out = {}
    model.train()
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    pred, att_t, att_s = model(datas, device="cpu", if_att=True)
    loss = (pred * w).sum()
    gr = torch.autograd.grad(loss, [p for p in model.parameters()], allow_unused=True)
    out = dict(state=state0, pred=pred.detach(), loss=loss.detach(), att_t=att_t.detach(), att_s=att_s.detach(),
               grads={n_: t for (n_, _), t in zip(model.named_parameters(), gr)})
    return dict(datas=saved, ctor=ctor, w=w, **out)


def main():
    torch.set_num_threads(1)
    rng = np.random.default_rng(7)
    out = {"tsp": tsp_case(rng)}
    out["cifar"] = attpool_case(rng, RM.HL_HGCNN_CIFAR10SP_dense_int3_attpool,
                                dict(channels=[1, 2, 1], filters=[8, 12, 16], mlp_channels=[10], K=3, node_dim=3, edge_dim=2,
                                     num_classes=3, keig=2, pool_loc=1, l=0.5), 31)
    out["pepfunc"] = attpool_case(rng, load_pepfunc_class(),
                                  dict(channels=[1, 2, 1], filters=[8, 12, 16], mlp_channels=[10], K=3, node_dim=3, edge_dim=2,
                                       num_classes=3, keig=2, pool_loc=1), 41)
    path = os.path.join(HERE, "models.pt")
    torch.save(out, path)
    print("models.pt", os.path.getsize(path))


if __name__ == "__main__":
    main()
