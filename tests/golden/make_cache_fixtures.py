"""Cache files in the reference's on-disk format (build container only: runs the reference's own construction).

  python tests/golden/make_cache_fixtures.py      # writes tests/golden/cache/*.pt

The reference writes `torch.save({'graph': PairData | [PairData, ...], 'maxeig': ..., 'par1': ...})`
(lib/Hodge_Dataset.py:475-476, :528-529) where `PairData` subclasses `torch_geometric.data.Data`.  torch_geometric is
not installable here, so its PICKLE LAYOUT is emulated: classes registered under the real module paths
(`torch_geometric.data.data.Data`, `torch_geometric.data.storage.GlobalStorage`, `lib.Hodge_Dataset.PairData`) whose
`__getstate__` reproduces what those classes pickle --
  * PyG >= 2.0:  Data.__dict__ = {'_store': GlobalStorage};  GlobalStorage state = {'_mapping': {attr: value},
    '_parent': <the Data object>} (its __getstate__ dereferences the weak parent link);
  * PyG 1.x:     Data.__dict__ = {attr: value}.
The graph CONTENT comes from the unmodified reference code: `process()` statement by statement (:447-474) with the
reference's own `adj2par1`, `eig_pe` (scipy eigh) and, for the multi-level files, its own `MLGC` (:517-529).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import lib.Hodge_Dataset as RD  # noqa: E402
from torch_geometric.utils import to_undirected, dense_to_sparse  # noqa: E402  (shim)
from make_golden import rand_graph  # noqa: E402


def reference_process(ei_dir, n, rng):
    """lib/Hodge_Dataset.py:447-474 on one raw graph (ZINC-style one-hot features + full eigenvector encodings)."""
    atom = torch.from_numpy(rng.integers(0, 21, n))
    bond = torch.from_numpy(rng.integers(1, 4, ei_dir.shape[1]))
    edge_index, edge_attr = to_undirected(ei_dir, bond, reduce='min')
    idx = edge_index[0] < edge_index[1]
    edge_index, edge_attr = edge_index[:, idx], edge_attr[idx]
    par1 = RD.adj2par1(edge_index, n, edge_index.shape[1]).to_dense()
    L0 = torch.matmul(par1, par1.T)
    lambda0, _ = torch.linalg.eigh(L0)
    maxeig = lambda0.max()
    L0 = 2 * torch.matmul(par1, par1.T) / maxeig
    L1 = 2 * torch.matmul(par1.T, par1) / maxeig
    node_pe = RD.eig_pe(L0, k=100)
    edge_pe = RD.eig_pe(L1, k=100)
    x_s = torch.nn.functional.one_hot(edge_attr - 1, num_classes=3)
    x_t = torch.nn.functional.one_hot(atom, num_classes=21)
    x_s = torch.cat([x_s.to(torch.float), edge_pe], dim=-1)
    x_t = torch.cat([x_t.to(torch.float), node_pe], dim=-1)
    data = RD.PairData(x_s=x_s, edge_index_s=None, edge_weight_s=None, x_t=x_t, edge_index_t=None, edge_weight_t=None,
                       y=torch.randn(1))
    data.edge_index_t, data.edge_weight_t = dense_to_sparse(L0)
    data.edge_index_s, data.edge_weight_s = dense_to_sparse(L1)
    data.num_node1 = data.x_t.shape[0]
    data.num_edge1 = data.x_s.shape[0]
    data.num_nodes = data.x_t.shape[0]
    data.edge_index = edge_index
    return data, maxeig, par1


# ---- the pickle layouts of torch_geometric, under the real module paths -------------------------------------------
def _install_fake_pyg():
    for name in ("torch_geometric.data.data", "torch_geometric.data.storage"):
        sys.modules[name] = types.ModuleType(name)

    class BaseStorage:
        def __init__(self, mapping, parent):
            self._mapping, self._parent = mapping, parent

        def __getstate__(self):
            return dict(self.__dict__)                         # `_parent` already dereferenced, as PyG's __getstate__ does

    class GlobalStorage(BaseStorage):
        pass

    class Data:
        pass

    for cls, mod in ((BaseStorage, "torch_geometric.data.storage"), (GlobalStorage, "torch_geometric.data.storage"),
                     (Data, "torch_geometric.data.data")):
        cls.__module__, cls.__qualname__ = mod, cls.__name__
        setattr(sys.modules[mod], cls.__name__, cls)

    class PairData(Data):
        pass

    PairData.__module__, PairData.__qualname__ = "lib.Hodge_Dataset", "PairData"
    real = sys.modules["lib.Hodge_Dataset"].PairData
    sys.modules["lib.Hodge_Dataset"].PairData = PairData       # pickle looks the class up by module path
    return PairData, GlobalStorage, real


def as_pyg2(data, PairData, GlobalStorage):
    attrs = {k: data[k] for k in data.keys}
    attrs["num_nodes"] = data.num_nodes
    out = PairData.__new__(PairData)
    out.__dict__["_store"] = GlobalStorage(attrs, out)
    return out


def as_pyg1(data, PairData):
    out = PairData.__new__(PairData)
    for k in data.keys:
        out.__dict__[k] = data[k]
    out.__dict__["__num_nodes__"] = data.num_nodes
    return out


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(11)
    out_dir = os.path.join(HERE, "cache")
    os.makedirs(out_dir, exist_ok=True)
    singles, multis, truth = [], [], []
    for i in range(4):
        n = int(rng.integers(9, 16))
        data, maxeig, par1 = reference_process(rand_graph(rng, n, 2), n, rng)
        singles.append((data, maxeig, par1))
    for i in range(3):
        n = int(rng.integers(10, 15))
        data, maxeig, par1 = reference_process(rand_graph(rng, n, 3), n, rng)
        datas = [data]
        temp, c_node, c_edge = RD.MLGC(datas[0])              # lib/Hodge_Dataset.py:523-527 (num_pool = 1)
        datas[0].x_t = torch.cat([c_node, datas[0].x_t], dim=-1)
        datas[0].x_s = torch.cat([c_edge, datas[0].x_s], dim=-1)
        datas.append(temp)
        multis.append(datas)
    # the plain tensors, for the test to compare with (no foreign classes in this file)
    plain = dict(single=[{k: d[k] for k in d.keys} for d, _, _ in singles],
                 multi=[[{k: d[k] for k in d.keys} for d in datas] for datas in multis],
                 maxeig=[m for _, m, _ in singles])
    torch.save(plain, os.path.join(out_dir, "expected.pt"))
    PairData, GlobalStorage, _ = _install_fake_pyg()
    for i, (d, maxeig, par1) in enumerate(singles):
        conv = as_pyg2(d, PairData, GlobalStorage) if i % 2 == 0 else as_pyg1(d, PairData)     # both layouts
        torch.save({'graph': conv, 'maxeig': maxeig, 'par1': par1}, os.path.join(out_dir, f"ZINC_BM_alleig_{i + 1}.pt"))
    for i, datas in enumerate(multis):
        torch.save({'graph': [as_pyg2(d, PairData, GlobalStorage) for d in datas]}, os.path.join(out_dir, f"ZINC_BM_MLGC_{i + 1}.pt"))
    print(sorted(os.listdir(out_dir)), sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir)), "bytes")


if __name__ == "__main__":
    main()
