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


# ---------------------------------------------------------------------------------------------
# two-level batches for the attention-pooling callers (peptides-func, CIFAR10-superpixel)
# ---------------------------------------------------------------------------------------------
def coarsen(ei, n):
    """Host restatement of MLGC (lib/Hodge_Dataset.py:241-295) with a deterministic greedy matching in
    place of torch_cluster.graclus (which is randomised): every node, in id order, pairs with its first
    still-unmatched neighbour.  Returns (ei1 [2,E1] coarse edges (imin, imax) in FIRST-APPEARANCE order,
    n1, c_node [n] int cluster ids, c_edge [E] float coarse-edge ids with +inf for edges inside a cluster)."""
    e = ei.shape[1]
    nbrs = [[] for _ in range(n)]
    for a, b in zip(ei[0].tolist(), ei[1].tolist()):
        nbrs[a].append(b)
        nbrs[b].append(a)
    cluster = np.full(n, -1, dtype=np.int64)
    for u in range(n):
        if cluster[u] >= 0:
            continue
        cluster[u] = u
        for v in sorted(nbrs[u]):
            if cluster[v] < 0:
                cluster[v] = u
                break
    uniq, c_node = np.unique(cluster, return_inverse=True)
    n1 = int(uniq.shape[0])
    c_edge = np.full(e, np.inf, dtype=np.float32)
    seen, lo_l, hi_l = {}, [], []
    ca, cb = c_node[ei[0]], c_node[ei[1]]
    for i in range(e):
        a, b = int(ca[i]), int(cb[i])
        if a == b:
            continue
        lo, hi = (a, b) if a < b else (b, a)
        k = seen.get((lo, hi))
        if k is None:
            k = len(lo_l)
            seen[(lo, hi)] = k
            lo_l.append(lo)
            hi_l.append(hi)
        c_edge[i] = k
    ei1 = np.array([lo_l, hi_l], dtype=np.int64).reshape(2, -1)
    return ei1, n1, c_node.astype(np.int64), c_edge


def make_multilevel_batch(shape="peptides", batch_size=64, seed=0, node_dim=None, edge_dim=None, num_targets=10):
    """[level-0 batch, level-1 batch] as the reference's DataLoader yields for the attpool models
    (lib/Hodge_Dataset.py:866-870 + list collation): column 0 of the level-0 x_t / x_s holds the
    per-graph cluster id (+inf for edges inside a cluster); level-1 features are the all-ones placeholders
    of MLGC (:283-284)."""
    n_lo, n_hi, kind, param, nd, ed = SHAPES[shape]
    nd, ed = node_dim or nd, edge_dim or ed
    rng = np.random.default_rng(seed)
    keys = ("edge_index", "edge_index_t", "edge_index_s", "edge_weight_t", "edge_weight_s")
    cols = [{k: [] for k in keys} for _ in range(2)]
    sizes = [([], []), ([], [])]
    cn, ce = [], []
    off = [[0, 0], [0, 0]]
    for _ in range(batch_size):
        n = int(rng.integers(n_lo, n_hi + 1))
        ei = _tree_plus_chords(rng, n, param) if kind == "tree" else _knn_graph(rng, n, param)
        ei1, n1, c_node, c_edge = coarsen(ei, n)
        cn.append(c_node.astype(np.float32))
        ce.append(c_edge)
        for lvl, (e_idx, nn_) in enumerate(((ei, n), (ei1, n1))):
            g = simplex_graph(e_idx, nn_)
            cols[lvl]["edge_index"].append(g["edge_index"] + off[lvl][0])
            cols[lvl]["edge_index_t"].append(g["edge_index_t"] + off[lvl][0])
            cols[lvl]["edge_index_s"].append(g["edge_index_s"] + off[lvl][1])
            cols[lvl]["edge_weight_t"].append(g["edge_weight_t"])
            cols[lvl]["edge_weight_s"].append(g["edge_weight_s"])
            sizes[lvl][0].append(nn_)
            sizes[lvl][1].append(e_idx.shape[1])
            off[lvl][0] += nn_
            off[lvl][1] += e_idx.shape[1]
    gen = torch.Generator().manual_seed(seed)
    out = []
    for lvl in range(2):
        b = SimpleNamespace()
        for k, v in cols[lvl].items():
            setattr(b, k, torch.from_numpy(np.ascontiguousarray(np.concatenate(v, axis=-1))))
        b.num_node1, b.num_edge1 = torch.tensor(sizes[lvl][0]), torch.tensor(sizes[lvl][1])
        b.num_graphs = batch_size
        if lvl == 0:
            b.x_t = torch.cat([torch.from_numpy(np.concatenate(cn)).view(-1, 1), torch.randn(off[0][0], nd, generator=gen)], -1)
            b.x_s = torch.cat([torch.from_numpy(np.concatenate(ce)).view(-1, 1), torch.randn(off[0][1], ed, generator=gen)], -1)
            b.y = torch.randn(batch_size, num_targets, generator=gen)
        else:
            b.x_t, b.x_s = torch.ones(off[1][0], 1), torch.ones(off[1][1], 1)
            b.y = torch.zeros(batch_size, 0)
        out.append(b)
    return out


def make_tsp_batch(batch_size=32, seed=0, n=500, k=25):
    """TSP-shaped batch (lib/Hodge_ST_Model.py:824-827): x_t = 2-D coordinates, x_s = [edge length, edge mask
    column of ones]; y = a Bernoulli label per edge."""
    SHAPES["_tsp"] = (n, n, "knn", k, 2, 1)
    b = make_batch("_tsp", batch_size, seed=seed, node_dim=2, edge_dim=1)
    gen = torch.Generator().manual_seed(seed + 1)
    b.x_t = torch.rand(b.x_t.shape[0], 2, generator=gen)
    length = (b.x_t[b.edge_index[0]] - b.x_t[b.edge_index[1]]).norm(dim=1, keepdim=True)
    b.x_s = torch.cat([length, torch.ones_like(length)], -1)
    b.y = (torch.rand(b.x_s.shape[0], generator=gen) < 0.15).long()
    return b
