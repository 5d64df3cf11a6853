"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules.

Runs only in the build container: needs /root/reference (read-only) and puts
`oracle/pyg_shim` (pure-torch stand-ins for torch_geometric / torch_scatter / torch_sparse /
torch_cluster, none installable here) on sys.path so `import lib.Hodge_Cheb_Conv` resolves.
The control flow that produces every number below is the reference's own
(lib/Hodge_Cheb_Conv.py, lib/Hodge_Dataset.py, lib/Hodge_ST_Model.py,
HL-HGAT-DEMO/lib/Hodge_Cheb_Conv.py); only the third-party primitives are shimmed.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.pt (deterministic)
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
sys.path.insert(0, "/root/reference")

import lib.Hodge_Cheb_Conv as RC  # noqa: E402
import lib.Hodge_Dataset as RD  # noqa: E402
import lib.Hodge_ST_Model as RM  # noqa: E402
from torch_geometric.data import Batch  # noqa: E402  (shim)
from torch_geometric.utils import to_undirected, dense_to_sparse, degree  # noqa: E402  (shim)


def rand_graph(rng, n, extra):
    """random tree + `extra` chords as a directed (both directions, shuffled) edge list."""
    und = set()
    for v in range(1, n):
        und.add((int(rng.integers(max(0, v - 3), v)), v))
    while len(und) < n - 1 + extra:
        a, b = sorted(int(t) for t in rng.integers(0, n, 2))
        if a != b:
            und.add((a, b))
    und = np.array(sorted(und)).T
    ei = np.concatenate([und, und[::-1]], 1)
    ei = ei[:, rng.permutation(ei.shape[1])]
    return torch.from_numpy(ei.copy()).long()


def ref_construct(ei_dir, n):
    """The reference's per-graph construction, statement by statement
    (lib/Hodge_Dataset.py:447-456, 467-468), on reference/shim functions."""
    attr = torch.arange(ei_dir.shape[1]) % 3 + 1
    edge_index, edge_attr = to_undirected(ei_dir, attr, reduce='min')
    idx = edge_index[0] < edge_index[1]
    edge_index, edge_attr = edge_index[:, idx], edge_attr[idx]
    par1 = RD.adj2par1(edge_index, n, edge_index.shape[1]).to_dense()
    L0 = torch.matmul(par1, par1.T)
    lambda0, _ = torch.linalg.eigh(L0)
    maxeig = lambda0.max()
    L0 = 2 * torch.matmul(par1, par1.T) / maxeig
    L1 = 2 * torch.matmul(par1.T, par1) / maxeig
    edge_index_t, edge_weight_t = dense_to_sparse(L0)
    edge_index_s, edge_weight_s = dense_to_sparse(L1)
    return dict(ei_dir=ei_dir, n=n, edge_index=edge_index, edge_attr=edge_attr, maxeig=maxeig,
                edge_index_t=edge_index_t, edge_weight_t=edge_weight_t,
                edge_index_s=edge_index_s, edge_weight_s=edge_weight_s)


def grads_of(module, out_sum_weights, out, inputs):
    loss = (out * out_sum_weights).sum()
    params = [p for p in module.parameters()]
    g = torch.autograd.grad(loss, list(inputs) + params, allow_unused=True)
    gi = [t.clone() if t is not None else None for t in g[:len(inputs)]]
    gp = {n: (t.clone() if t is not None else None)
          for (n, _), t in zip(module.named_parameters(), g[len(inputs):])}
    return gi, gp


def conv_cases(rng):
    out = []
    g = ref_construct(rand_graph(rng, 11, 3), 11)
    for fam, cls in (("laguerre", RC.HodgeLaguerreConv), ("cheb", RC.HodgeChebConv)):
        for K in (1, 2, 3, 5):
            for side, fin, fout, three_d in (("t", 6, 5, False), ("s", 4, 8, False), ("t", 3, 4, True)):
                if three_d and fam == "cheb" and K > 1:
                    # reference bug: lib/Hodge_Cheb_Conv.py:410-411 calls .view on a transposed
                    # tensor -> RuntimeError; the 3-D Chebyshev path cannot run in the reference
                    continue
                torch.manual_seed(100 * K + fin)
                conv = cls(fin, fout, K)
                with torch.no_grad():
                    conv.bias.normal_()
                ei, ew = g[f"edge_index_{side}"], g[f"edge_weight_{side}"]
                r = g["n"] if side == "t" else g["edge_index"].shape[1]
                x = torch.randn((r, 3, fin) if three_d else (r, fin), requires_grad=True)
                y = conv(x, ei, ew)
                wsum = torch.randn_like(y)
                gi, gp = grads_of(conv, wsum, y, [x])
                out.append(dict(family=fam, K=K, side=side, fin=fin, fout=fout, three_d=three_d, x=x.detach(),
                                edge_index=ei, edge_weight=ew, state=conv.state_dict(),
                                y=y.detach(), wsum=wsum, gx=gi[0], gp=gp))
    return dict(graph=g, cases=out)


def fastconv_cases(rng):
    spec = importlib.util.spec_from_file_location(
        "demo_conv", "/root/reference/HL-HGAT-DEMO/lib/Hodge_Cheb_Conv.py")
    # the DEMO fork does `from lib.Hodge_Dataset import *`: point `lib` at the DEMO tree
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "lib" or k.startswith("lib.")}
    sys.path.insert(0, "/root/reference/HL-HGAT-DEMO")
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "mat73", "networkx"):  # plotting-only imports
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    try:
        demo = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(demo)
    finally:
        sys.path.pop(0)
        for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    import torch_sparse
    g = ref_construct(rand_graph(rng, 9, 2), 9)
    out = []
    for K in (2, 3, 4):
        torch.manual_seed(K)
        conv = demo.HodgeLaguerreFastConv(5, 6, K)
        ei, ew = g["edge_index_s"], g["edge_weight_s"]
        adj_t = torch_sparse.SparseTensor(row=ei[0], col=ei[1], value=ew,
                                          sparse_sizes=(g["edge_index"].shape[1],) * 2).t()
        x = torch.randn(g["edge_index"].shape[1], 5)
        out.append(dict(K=K, x=x, edge_index=ei, edge_weight=ew, state=conv.state_dict(),
                        y=conv(x, adj_t).detach()))
    return out


def neint_cases(rng):
    g = ref_construct(rand_graph(rng, 13, 4), 13)
    n, e = g["n"], g["edge_index"].shape[1]
    par = RD.adj2par1(g["edge_index"], n, e)
    D = degree(g["edge_index"].view(-1), num_nodes=n) + 1e-6
    out = []
    for only_att, sig in ((False, "sigmoid"), (True, "sigmoid"), (True, "relu")):
        torch.manual_seed(7)
        mod = RC.NodeEdgeInt(d=6, dk=4, dv=5, only_att=only_att,
                             sigma=torch.nn.Sigmoid() if sig == "sigmoid" else torch.nn.ReLU(), l=0.5)
        mod.train()
        x_t = torch.randn(n, 6, requires_grad=True)
        x_s = torch.randn(e, 6, requires_grad=True)
        y_t, y_s = mod(x_t, x_s, par, D)
        w_t, w_s = torch.randn_like(y_t), torch.randn_like(y_s)
        loss = (y_t * w_t).sum() + (y_s * w_s).sum()
        params = list(mod.parameters())
        gr = torch.autograd.grad(loss, [x_t, x_s] + params)
        out.append(dict(only_att=only_att, sigma=sig, l=0.5, d=6, dk=4, dv=5,
                        x_t=x_t.detach(), x_s=x_s.detach(), D=D, edge_index=g["edge_index"],
                        state={k: v.clone() for k, v in mod.state_dict().items()
                               if "running" not in k and "num_batches" not in k},
                        y_t=y_t.detach(), y_s=y_s.detach(), w_t=w_t, w_s=w_s,
                        gx_t=gr[0], gx_s=gr[1],
                        gp={n_: t for (n_, _), t in zip(mod.named_parameters(), gr[2:])}))
    # raw transfers on the SURVEY tiny graph
    ei = torch.tensor([[0, 0, 1, 2], [1, 2, 2, 3]])
    par = RD.adj2par1(ei, 4, 4)
    Dt = degree(ei.view(-1), num_nodes=4) + 1e-6
    x_s = torch.tensor([[1.], [2.], [3.], [4.]])
    x_t = torch.tensor([[0.], [10.], [20.], [30.]])
    tiny = dict(edge_index=ei, par_dense=par.to_dense(), D=Dt,
                x_s2t=(1 / Dt).view(-1, 1) * torch.sparse.mm(par.abs(), x_s),
                x_t2s=torch.sparse.mm(par.abs().transpose(0, 1), x_t) / 2)
    return dict(graph=g, cases=out, tiny=tiny)


def pool_case(rng):
    """MLGC (lib/Hodge_Dataset.py:241-295, graclus shimmed deterministically) + SAPool
    (lib/Hodge_Cheb_Conv.py:36-59)."""
    torch.manual_seed(11)
    g = ref_construct(rand_graph(rng, 14, 4), 14)
    n, e = g["n"], g["edge_index"].shape[1]
    fine = RD.PairData(x_s=torch.randn(e, 6), edge_index_s=g["edge_index_s"], edge_weight_s=g["edge_weight_s"],
                       x_t=torch.randn(n, 6), edge_index_t=g["edge_index_t"], edge_weight_t=g["edge_weight_t"])
    fine.edge_index = g["edge_index"]
    fine.num_node1, fine.num_edge1, fine.num_nodes = n, e, n
    coarse, c_node, c_edge = RD.MLGC(fine)
    par = RD.adj2par1(g["edge_index"], n, e)
    D = degree(g["edge_index"].view(-1), num_nodes=n) + 1e-6
    pool = RC.SAPool(d=6, dk=4)
    res = pool(fine.x_t, fine.x_s, par, D, [fine, coarse], [c_node.float()], [c_edge], 0, device='cpu')
    return dict(fine=g, x_t=fine.x_t, x_s=fine.x_s, D=D, c_node=c_node, c_edge=c_edge,
                coarse=dict(edge_index=coarse.edge_index, edge_index_t=coarse.edge_index_t,
                            edge_weight_t=coarse.edge_weight_t, edge_index_s=coarse.edge_index_s,
                            edge_weight_s=coarse.edge_weight_s, n=coarse.num_node1, e=coarse.num_edge1),
                state=pool.state_dict(), x_t1=res[0].detach(), x_s1=res[1].detach(),
                par1_dense=res[2].to_dense(), D1=res[3], att_t=res[9].detach(), att_s=res[10].detach())


def zinc_model_case(rng):
    """HL_HGCNN_zinc_dense_int3_pyr forward + backward on a collated batch of 6 graphs
    (lib/Hodge_ST_Model.py:544-646; collation lib/Hodge_Dataset.py:40-48)."""
    torch.manual_seed(5)
    graphs, raw = [], []
    for _ in range(6):
        n = int(rng.integers(8, 14))
        g = ref_construct(rand_graph(rng, n, 2), n)
        e = g["edge_index"].shape[1]
        d = RD.PairData(x_s=torch.randn(e, 5), edge_index_s=g["edge_index_s"], edge_weight_s=g["edge_weight_s"],
                        x_t=torch.randn(n, 7), edge_index_t=g["edge_index_t"], edge_weight_t=g["edge_weight_t"],
                        y=torch.randn(1))
        d.num_node1, d.num_edge1, d.num_nodes, d.edge_index = n, e, n, g["edge_index"]
        graphs.append(d)
        raw.append(dict(ei_dir=g["ei_dir"], n=n))
    batch = Batch.from_data_list(graphs)
    out = {}
    for K in (2, 3):
        torch.manual_seed(K)
        model = RM.HL_HGCNN_zinc_dense_int3_pyr(channels=[1, 2], filters=[8, 12], mlp_channels=[10], K=K,
                                                node_dim=4, edge_dim=2, keig=3)
        model.train()
        state0 = {k: v.clone() for k, v in model.state_dict().items()}
        pred = model(batch, device='cpu')
        loss = torch.nn.functional.l1_loss(pred, batch.y.view(-1, 1))
        gr = torch.autograd.grad(loss, [p for p in model.parameters()], allow_unused=True)
        out[K] = dict(state=state0, pred=pred.detach(), loss=loss.detach(),
                      grads={n_: t for (n_, _), t in zip(model.named_parameters(), gr)})
    keys = ("x_t", "x_s", "edge_index", "edge_index_t", "edge_index_s", "edge_weight_t",
            "edge_weight_s", "y", "num_node1", "num_edge1")
    return dict(batch={k: batch[k] for k in keys}, raw=raw, runs=out,
                ctor=dict(channels=[1, 2], filters=[8, 12], mlp_channels=[10], node_dim=4, edge_dim=2, keig=3))


def main():
    torch.set_num_threads(1)
    rng = np.random.default_rng(0)
    torch.save(conv_cases(rng), os.path.join(HERE, "conv.pt"))
    torch.save(fastconv_cases(rng), os.path.join(HERE, "fastconv.pt"))
    torch.save(neint_cases(rng), os.path.join(HERE, "neint.pt"))
    torch.save(pool_case(rng), os.path.join(HERE, "pool.pt"))
    torch.save(zinc_model_case(rng), os.path.join(HERE, "zinc_model.pt"))
    cons = [ref_construct(rand_graph(rng, n, x), n) for n, x in ((5, 1), (12, 0), (20, 6), (30, 12))]
    # an isolated node: node 6 has no edge (dense_to_sparse drops its all-zero L0 row)
    cons.append(ref_construct(torch.tensor([[0, 1, 1, 2, 3, 4, 4, 5], [1, 0, 2, 1, 4, 3, 5, 4]]), 7))
    torch.save(cons, os.path.join(HERE, "construct.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
