"""Synthetic mini-batches shaped like the BASELINE.json configs, built on
the host with numpy/scipy exactly per the reference's construction formulas
(lib/Hodge_Dataset.py:447-456,467-468: i<j edges in lexicographic order, L0 = 2 B1 B1^T / lmax,
L1 = 2 B1^T B1 / lmax, row-major COO of the nonzeros) and collated block-diagonally
(lib/Hodge_Dataset.py:40-48).  Data preparation only -- not on the timed path."""
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp
import torch


def _tree_plus_chords(rng, n, chords, back=3):
    und = set()
    for v in range(1, n):
        und.add((int(rng.integers(max(0, v - back), v)), v))
    target = min(n - 1 + chords, n * (n - 1) // 2)
    while len(und) < target:
        a, b = sorted(int(t) for t in rng.integers(0, n, 2))
        if a != b:
            und.add((a, b))
    return np.array(sorted(und), dtype=np.int64).T          # [2,E], i<j, lexicographic


def _knn_graph(rng, n, k):
    pts = rng.random((n, 2))
    d = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d, np.inf)
    nb = np.argsort(d, 1)[:, :k]
    a = np.repeat(np.arange(n), k)
    b = nb.reshape(-1)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = np.unique(lo * n + hi)
    return np.stack([key // n, key % n]).astype(np.int64)


def simplex_graph(ei, n):
    """ei: [2,E] undirected i<j lexicographic.  Returns the reference's per-graph tensors."""
    e = ei.shape[1]
    ar = np.arange(e)
    b1 = sp.csr_matrix((np.concatenate([-np.ones(e), np.ones(e)]), (np.concatenate([ei[0], ei[1]]), np.concatenate([ar, ar]))),
                       shape=(n, e))
    l0 = (b1 @ b1.T).tocsr()
    l1 = (b1.T @ b1).tocsr()
    lmax = np.float32(np.linalg.eigvalsh(l0.toarray().astype(np.float32)).max())
    out = {}
    for name, m in (("t", l0), ("s", l1)):
        m.sort_indices()
        m.eliminate_zeros()
        coo = m.tocoo()                                    # row-major, ascending columns (= dense nonzero())
        out["edge_index_" + name] = np.stack([coo.row, coo.col]).astype(np.int64)
        out["edge_weight_" + name] = (np.float32(2.0) * coo.data.astype(np.float32)) / lmax
    out["edge_index"] = ei
    out["lmax"] = lmax
    return out


SHAPES = {
    # name: (n_lo, n_hi, kind, param, node_feat, edge_feat)
    "zinc": (18, 28, "tree", 3, 28, 10),
    "peptides": (120, 180, "tree", 4, 19, 13),
    "cifar": (110, 125, "knn", 8, 15, 14),
    "tsp": (500, 500, "knn", 25, 2, 1),
}


def make_batch(shape="zinc", batch_size=128, seed=0, node_dim=None, edge_dim=None, num_targets=1):
    """A collated batch (CPU tensors) with the attribute names of the reference's PairData batch."""
    n_lo, n_hi, kind, param, nd, ed = SHAPES[shape]
    nd, ed = node_dim or nd, edge_dim or ed
    rng = np.random.default_rng(seed)
    cols = {k: [] for k in ("edge_index", "edge_index_t", "edge_index_s", "edge_weight_t", "edge_weight_s")}
    nn1, ne1 = [], []
    n_off = e_off = 0
    for _ in range(batch_size):
        n = int(rng.integers(n_lo, n_hi + 1))
        ei = _tree_plus_chords(rng, n, param) if kind == "tree" else _knn_graph(rng, n, param)
        g = simplex_graph(ei, n)
        e = ei.shape[1]
        cols["edge_index"].append(g["edge_index"] + n_off)
        cols["edge_index_t"].append(g["edge_index_t"] + n_off)
        cols["edge_index_s"].append(g["edge_index_s"] + e_off)
        cols["edge_weight_t"].append(g["edge_weight_t"])
        cols["edge_weight_s"].append(g["edge_weight_s"])
        nn1.append(n)
        ne1.append(e)
        n_off += n
        e_off += e
    gen = torch.Generator().manual_seed(seed)
    b = SimpleNamespace()
    for k, v in cols.items():
        arr = np.ascontiguousarray(np.concatenate(v, axis=-1))
        setattr(b, k, torch.from_numpy(arr))
    b.x_t = torch.randn(n_off, nd, generator=gen)
    b.x_s = torch.randn(e_off, ed, generator=gen)
    b.y = torch.randn(batch_size, num_targets, generator=gen)
    b.num_node1 = torch.tensor(nn1)
    b.num_edge1 = torch.tensor(ne1)
    b.num_graphs = batch_size
    return b


TENSOR_KEYS = ("x_t", "x_s", "y", "edge_index", "edge_index_t", "edge_index_s", "edge_weight_t", "edge_weight_s",
               "num_node1", "num_edge1")


def batch_to(b, device, non_blocking=False):
    out = SimpleNamespace(num_graphs=b.num_graphs)
    for k in TENSOR_KEYS:
        setattr(out, k, getattr(b, k).to(device, non_blocking=non_blocking))
    return out


def pin_batch(b):
    out = SimpleNamespace(num_graphs=b.num_graphs)
    for k in TENSOR_KEYS:
        setattr(out, k, getattr(b, k).pin_memory())
    return out


def batch_nbytes(b):
    return sum(getattr(b, k).numel() * getattr(b, k).element_size() for k in TENSOR_KEYS)
