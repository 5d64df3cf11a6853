"""Semantics restated from the PyG documentation."""
import torch
from .num_nodes import maybe_num_nodes


def degree(index, num_nodes=None, dtype=None):
    n = maybe_num_nodes(index, num_nodes)
    out = torch.zeros((n,), dtype=dtype or torch.get_default_dtype(), device=index.device)
    one = torch.ones((index.size(0),), dtype=out.dtype, device=index.device)
    return out.scatter_add_(0, index, one)


def coalesce(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    n = maybe_num_nodes(edge_index, num_nodes)
    key = edge_index[0] * n + edge_index[1]
    key_sorted, perm = torch.sort(key, stable=True)
    ei = edge_index[:, perm]
    uniq, inv = torch.unique_consecutive(key_sorted, return_inverse=True)
    first = torch.ones_like(key_sorted, dtype=torch.bool)
    first[1:] = key_sorted[1:] != key_sorted[:-1]
    ei_out = ei[:, first]
    if edge_attr is None:
        return ei_out
    ea = edge_attr[perm]
    shape = (uniq.numel(),) + tuple(ea.shape[1:])
    idx = inv.view((-1,) + (1,) * (ea.dim() - 1)).expand_as(ea)
    if reduce in ("add", "sum"):
        out = torch.zeros(shape, dtype=ea.dtype).scatter_add_(0, idx, ea)
    elif reduce == "mean":
        out = torch.zeros(shape, dtype=ea.dtype).scatter_add_(0, idx, ea)
        cnt = torch.zeros((uniq.numel(),), dtype=ea.dtype).scatter_add_(
            0, inv, torch.ones_like(inv, dtype=ea.dtype))
        out = out / cnt.view((-1,) + (1,) * (ea.dim() - 1))
    elif reduce == "min":
        out = torch.zeros(shape, dtype=ea.dtype).scatter_reduce_(0, idx, ea, "amin", include_self=False)
    elif reduce == "max":
        out = torch.zeros(shape, dtype=ea.dtype).scatter_reduce_(0, idx, ea, "amax", include_self=False)
    else:
        raise ValueError(reduce)
    return ei_out, out


def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    row, col = edge_index[0], edge_index[1]
    ei = torch.stack([torch.cat([row, col]), torch.cat([col, row])], dim=0)
    if edge_attr is None:
        return coalesce(ei, None, num_nodes, reduce)
    ea = torch.cat([edge_attr, edge_attr], dim=0)
    return coalesce(ei, ea, num_nodes, reduce)


def dense_to_sparse(adj):
    assert adj.dim() == 2
    idx = adj.nonzero().t().contiguous()
    return idx, adj[idx[0], idx[1]]


def add_self_loops(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
    n = maybe_num_nodes(edge_index, num_nodes)
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    ei = torch.cat([edge_index, torch.stack([loop, loop])], dim=1)
    if edge_attr is not None:
        fill = edge_attr.new_full((n,) + tuple(edge_attr.shape[1:]), fill_value)
        edge_attr = torch.cat([edge_attr, fill], dim=0)
    return ei, edge_attr


def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None,
             return_edge_mask=False):
    n = maybe_num_nodes(edge_index, num_nodes)
    if subset.dtype == torch.bool:
        node_mask = subset
    else:
        node_mask = torch.zeros(n, dtype=torch.bool)
        node_mask[subset] = True
    edge_mask = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    ei = edge_index[:, edge_mask]
    ea = edge_attr[edge_mask] if edge_attr is not None else None
    if relabel_nodes:
        remap = torch.zeros(n, dtype=torch.long)
        remap[node_mask] = torch.arange(int(node_mask.sum()))
        ei = remap[ei]
    if return_edge_mask:
        return ei, ea, edge_mask
    return ei, ea


def to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    import scipy.sparse as sp
    n = maybe_num_nodes(edge_index, num_nodes)
    v = edge_attr if edge_attr is not None else torch.ones(edge_index.size(1))
    return sp.coo_matrix((v.numpy(), (edge_index[0].numpy(), edge_index[1].numpy())), (n, n))


def remove_isolated_nodes(edge_index, edge_attr=None, num_nodes=None):
    n = maybe_num_nodes(edge_index, num_nodes)
    mask = torch.zeros(n, dtype=torch.bool)
    mask[edge_index.view(-1)] = True
    remap = torch.full((n,), -1, dtype=torch.long)
    remap[mask] = torch.arange(int(mask.sum()))
    return remap[edge_index], edge_attr, mask


def unbatch(src, batch, dim=0):
    sizes = degree(batch, dtype=torch.long).tolist()
    return src.split(sizes, dim)


def unbatch_edge_index(edge_index, batch):
    deg = degree(batch, dtype=torch.int64)
    ptr = torch.cat([deg.new_zeros(1), deg.cumsum(dim=0)[:-1]], dim=0)
    edge_batch = batch[edge_index[0]]
    edge_index = edge_index - ptr[edge_batch]
    sizes = degree(edge_batch, dtype=torch.int64).cpu().tolist()
    return edge_index.split(sizes, dim=1)


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    n = maybe_num_nodes(index, num_nodes)
    mx = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype)
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    mx = mx.scatter_reduce_(0, idx, src, "amax", include_self=True)
    out = (src - mx[index]).exp()
    s = torch.zeros_like(mx).scatter_add_(0, idx, out)
    return out / (s[index] + 1e-16)
