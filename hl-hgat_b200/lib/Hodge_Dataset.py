"""Host-side mirror of the hot-path pieces of the reference's lib/Hodge_Dataset.py: the `PairData` container
with its block-diagonal increments (:27-48), the mini-batch collation PyG's DataLoader performs with it
(including the list-of-levels samples of the attention-pooling datasets, :515 / :650), `adj2par1` (:169-191)
and the simplex-graph construction of `Dataset.process` / `get` (:442-477) -- here done for a WHOLE mini-batch
on the GPU by `construct.build_simplex_batch`, never forming a dense N x E matrix.

No torch_geometric: `PairData` is a plain attribute container, `collate` is pure torch (host data
preparation, like the reference's DataLoader workers)."""
from types import SimpleNamespace

import torch

from ..coarsen import mlgc_batch  # noqa: F401  (MLGC / MLGC_weighted :241-353 for a whole mini-batch, on the GPU)
from ..construct import build_simplex_batch
from .Hodge_Cheb_Conv import adj2par1  # noqa: F401  (same name / signature as lib/Hodge_Dataset.py:169)

__all__ = ["PairData", "collate", "adj2par1", "simplex_batch_from_graphs", "mlgc_batch", "two_level_batch_from_graphs"]


class PairData(SimpleNamespace):
    """Reference lib/Hodge_Dataset.py:27-48.  Same constructor keywords; further attributes (`num_node1`,
    `num_edge1`, `num_nodes`, ...) are set by assignment exactly like the reference does (:470-474)."""

    def __init__(self, edge_index_s=None, x_s=None, edge_index_t=None, x_t=None, edge_weight_s=None, edge_weight_t=None,
                 edge_index=None, y=None):
        super().__init__(edge_index_s=edge_index_s, x_s=x_s, edge_index_t=edge_index_t, x_t=x_t,
                         edge_weight_s=edge_weight_s, edge_weight_t=edge_weight_t, edge_index=edge_index, y=y)

    def __inc__(self, key, value=None, *args, **kwargs):
        """Offsets added to index tensors when graphs are stacked block-diagonally (:40-48)."""
        if key == "edge_index_s":
            return self.x_s.size(0)
        if key in ("edge_index", "edge_index_t"):
            return self.x_t.size(0)
        if "index" in key:                                   # PyG default for other *index* keys: num_nodes
            return getattr(self, "num_nodes", None) or self.x_t.size(0)
        return 0

    def __cat_dim__(self, key, value=None, *args, **kwargs):
        return -1 if "index" in key else 0

    def to(self, device, non_blocking=False):
        for k, v in vars(self).items():
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self


def _collate_level(graphs):
    first = graphs[0]
    keys = [k for k, v in vars(first).items() if v is not None and k != "num_nodes"]
    out = PairData()
    inc = {k: 0 for k in keys}
    cols = {k: [] for k in keys}
    for g in graphs:
        for k in keys:
            v = getattr(g, k)
            if torch.is_tensor(v) and v.dim() > 0:
                cols[k].append(v + inc[k] if inc[k] else v)
                inc[k] += int(g.__inc__(k, v))
            else:
                cols[k].append(v)
    for k in keys:
        v0 = cols[k][0]
        if torch.is_tensor(v0) and v0.dim() > 0:
            setattr(out, k, torch.cat(cols[k], dim=first.__cat_dim__(k, v0)))
        elif torch.is_tensor(v0):
            setattr(out, k, torch.stack(cols[k]))
        elif isinstance(v0, (int, float)):
            setattr(out, k, torch.tensor(cols[k]))          # python numbers collate to a [B] tensor
        else:
            setattr(out, k, cols[k])
    sizes = torch.tensor([g.x_t.size(0) for g in graphs])
    out.batch = torch.repeat_interleave(torch.arange(len(graphs)), sizes)
    out.ptr = torch.cat([sizes.new_zeros(1), sizes.cumsum(0)])
    out.num_nodes = int(sizes.sum())
    out.num_graphs = len(graphs)
    return out


def collate(samples):
    """What `torch_geometric.loader.DataLoader` yields for a list of samples: a block-diagonal batch for
    `PairData` samples, and for samples that are LISTS of `PairData` (fine graph + coarsened levels) the
    list of per-level batches `[Batch(level 0), Batch(level 1), ...]`."""
    if isinstance(samples[0], (list, tuple)):
        return [_collate_level([s[l] for s in samples]) for l in range(len(samples[0]))]
    return _collate_level(samples)


def simplex_batch_from_graphs(edge_indices, num_nodes, device="cuda:0", edge_attrs=None):
    """`Dataset.process` (lib/Hodge_Dataset.py:447-456,467-468) for a list of raw graphs at once: directed (or
    undirected, any order, duplicates allowed) `edge_index` tensors [2, M_g] with LOCAL node ids and the node
    counts.  Returns the device-resident `construct.SimplexBatch` (B1 tables, L0 / L1 in block-diagonal CSR
    scaled by 2 / lambda_max per graph, bit-exact indices and signs)."""
    counts = torch.as_tensor(list(num_nodes), dtype=torch.int64)
    offs = torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]])
    src = torch.cat([ei[0] + o for ei, o in zip(edge_indices, offs)])
    dst = torch.cat([ei[1] + o for ei, o in zip(edge_indices, offs)])
    attr = None if edge_attrs is None else torch.cat(list(edge_attrs)).to(device)
    return build_simplex_batch(src.to(device), dst.to(device), counts, edge_attr=attr)


def _as_level(sb, x_t, x_s, y=None):
    """A `SimplexBatch` in the attribute layout the model classes read (the reference's collated PairData batch),
    with the device-side CSR / incidence tables attached so that nothing is re-bucketed."""
    d = PairData(x_t=x_t, x_s=x_s, edge_index=sb.edge_index, y=y)
    d.edge_index_t, d.edge_weight_t = sb.coo("t")
    d.edge_index_s, d.edge_weight_s = sb.coo("s")
    d.num_node1, d.num_edge1 = sb.num_node1, sb.num_edge1
    d.num_nodes, d.num_graphs = sb.num_nodes, sb.num_graphs
    d.op_t, d.op_s, d.incidence = sb.op_t, sb.op_s, sb.incidence
    return d


def two_level_batch_from_graphs(edge_indices, num_nodes, x_t, x_s, y=None, device="cuda:0", edge_weight=None):
    """The `[fine batch, coarse batch]` pair the attention-pooling models consume, built on the GPU from raw graphs:
    `Dataset.get` (lib/Hodge_Dataset.py:829-870) = construction + MLGC + cluster ids prepended as column 0 of the
    fine features (:867-868) + list collation.  `x_t` [sum N, F_t] / `x_s` [sum E, F_s] are the fine node / edge
    features (edges in the lexicographic i<j order of every graph); `edge_weight` (per fine edge) selects
    MLGC_weighted's heavy-edge matching."""
    sb0 = simplex_batch_from_graphs(edge_indices, num_nodes, device=device)
    sb1, c_node, c_edge = mlgc_batch(sb0, edge_weight)
    lv0 = _as_level(sb0, torch.cat([c_node, x_t.to(device)], -1), torch.cat([c_edge, x_s.to(device)], -1), y)
    lv1 = _as_level(sb1, torch.ones(sb1.num_nodes, 1, device=device), torch.ones(sb1.num_edges, 1, device=device))
    return [lv0, lv1]
