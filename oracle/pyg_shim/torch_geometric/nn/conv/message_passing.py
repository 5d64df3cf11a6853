import inspect
import torch


class MessagePassing(torch.nn.Module):
    """flow = source_to_target: x_j = x[edge_index[0]]; messages are materialised, then summed
    into rows edge_index[1] of a zero tensor with x.size(0) rows (aggr='add')."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kwargs):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target" and node_dim == 0
        self.aggr = aggr
        self._msg_params = [p for p in inspect.signature(self.message).parameters]

    def propagate(self, edge_index, size=None, **kwargs):
        n_out = None
        args = {}
        for name in self._msg_params:
            if name.endswith("_j"):
                src = kwargs[name[:-2]]
                args[name] = src.index_select(0, edge_index[0])
                n_out = src.size(0)
            elif name.endswith("_i"):
                src = kwargs[name[:-2]]
                args[name] = src.index_select(0, edge_index[1])
                n_out = src.size(0)
            else:
                args[name] = kwargs[name]
        if size is not None:
            n_out = size[1] if isinstance(size, (tuple, list)) else size
        msg = self.message(**args)
        out = torch.zeros((n_out,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        return out.index_add_(0, edge_index[1], msg)

    def message(self, x_j):  # pragma: no cover
        return x_j
