"""torch-cluster's graclus is a randomized greedy matching and is NOT reproducible
(SURVEY.md Appendix A); this stand-in is a *seeded-order deterministic* greedy heavy-edge
matching with the same output contract: cluster id per node = min member id of its pair."""
import torch


def graclus_cluster(row, col, weight=None, num_nodes=None):
    n = int(num_nodes) if num_nodes is not None else int(max(row.max(), col.max())) + 1
    row_l, col_l = row.tolist(), col.tolist()
    w = weight.tolist() if weight is not None else [1.0] * len(row_l)
    nbrs = [[] for _ in range(n)]
    for r, c, ww in zip(row_l, col_l, w):
        if r != c:
            nbrs[r].append((c, ww))
    cluster = [-1] * n
    for u in range(n):
        if cluster[u] >= 0:
            continue
        best, best_w = -1, float("-inf")
        for v, ww in nbrs[u]:
            if cluster[v] < 0 and ww > best_w:
                best, best_w = v, ww
        if best >= 0:
            cluster[u] = cluster[best] = min(u, best)
        else:
            cluster[u] = u
    return torch.tensor(cluster, dtype=torch.long)
