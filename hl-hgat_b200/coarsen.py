"""Multi-level graph coarsening (MLGC) for a whole mini-batch on the GPU.

Reference: `MLGC` / `MLGC_weighted`, lib/Hodge_Dataset.py:241-353 -- per graph, on the CPU: graclus matching,
an O(E) Python loop that numbers the coarse edges in first-appearance order, dense B1 / eigh / dense_to_sparse
for the coarse operators.  Here: one warp per graph for the (deterministic) greedy matching
(`hl_greedy_matching`), sort / scan primitives for the relabelling and the first-appearance numbering, and the
same sort/segment construction kernels as level 0 for the coarse L0 / L1 (`construct.simplex_batch_from_edges`).
"""
import torch

from . import _native as N
from .construct import SimplexBatch, simplex_batch_from_edges

INF = float("inf")


def greedy_matching(sb, edge_weight=None):
    """cluster[u] = id of the smaller member of u's pair (or u itself), int32 [N] with GLOBAL node ids."""
    dev = sb.tail.device
    node_ptr = torch.zeros(sb.num_graphs + 1, dtype=torch.int32, device=dev)
    node_ptr[1:] = torch.cumsum(sb.num_node1.to(dev), 0)
    cluster = torch.empty(sb.num_nodes, dtype=torch.int32, device=dev)
    w = None if edge_weight is None else edge_weight.to(dev, torch.float32).contiguous()
    inc = sb.incidence
    N.check(N.lib().hl_greedy_matching(node_ptr.data_ptr(), sb.num_graphs, inc.rowptr.data_ptr(), inc.edge.data_ptr(),
                                       sb.tail.data_ptr(), sb.head.data_ptr(), N.ptr(w), cluster.data_ptr(), N.stream_ptr()),
            "hl_greedy_matching")
    return cluster


def mlgc_batch(sb: SimplexBatch, edge_weight=None):
    """Coarsen every graph of `sb` once.  Returns `(coarse, c_node, c_edge)`:
    `coarse`  SimplexBatch of the level-1 graphs (coarse edges `(imin, imax)` in first-appearance order),
    `c_node`  float32 [N,1]: per-graph (local) cluster id of every fine node            (lib/Hodge_Dataset.py:295),
    `c_edge`  float32 [E,1]: per-graph (local) coarse-edge id of every fine edge, +inf for edges inside a cluster.
    `edge_weight` (per fine edge, e.g. exp(-x_s[:,0]**2) of MLGC_weighted :309) selects the heavy-edge variant."""
    dev = sb.tail.device
    n, e, G = sb.num_nodes, sb.num_edges, sb.num_graphs
    cluster = greedy_matching(sb, edge_weight).long()
    # cluster representatives in ascending id order -> consecutive coarse node ids (torch.unique + dict, :256-261)
    is_rep = cluster == torch.arange(n, device=dev)
    rep_rank = torch.cumsum(is_rep, 0) - 1                                  # global coarse id of a representative
    cid = rep_rank[cluster]                                                 # global coarse id of every fine node
    node_graph = sb.node_graph.long()
    n1 = torch.zeros(G, dtype=torch.int64, device=dev).index_add_(0, node_graph, is_rep.long())
    n1_off = torch.cumsum(n1, 0) - n1
    c_node = (cid - n1_off[node_graph]).to(torch.float32).view(-1, 1)
    # coarse edges: key (imin, imax); numbering by first appearance in the fine edge order (:262-275)
    ca, cb = cid[sb.tail.long()], cid[sb.head.long()]
    lo, hi = torch.minimum(ca, cb), torch.maximum(ca, cb)
    cross = lo != hi
    n1_tot = int(n1.sum())
    pos = torch.nonzero(cross).view(-1)
    key = lo[pos] * max(n1_tot, 1) + hi[pos]
    uniq, inv = torch.unique(key, return_inverse=True)
    first = torch.full((uniq.numel(),), e, dtype=torch.int64, device=dev).scatter_reduce_(0, inv, pos, "amin")
    order = torch.argsort(first)                                            # coarse edge r = the r-th key to appear
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=dev)
    tail1, head1 = (uniq[order] // max(n1_tot, 1)), (uniq[order] % max(n1_tot, 1))
    coarse_graph = torch.repeat_interleave(torch.arange(G, device=dev), n1, output_size=n1_tot)
    e1 = torch.zeros(G, dtype=torch.int64, device=dev).index_add_(0, coarse_graph[tail1], torch.ones_like(tail1))
    e1_off = torch.cumsum(e1, 0) - e1
    c_edge = torch.full((e,), INF, dtype=torch.float32, device=dev)
    c_edge[pos] = (rank[inv] - e1_off[node_graph[sb.tail.long()[pos]]]).to(torch.float32)
    coarse = simplex_batch_from_edges(tail1, head1, n1)
    return coarse, c_node, c_edge.view(-1, 1)
