import torch


def maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return int(num_nodes)
    if edge_index.numel() == 0:
        return 0
    return int(edge_index.max()) + 1
