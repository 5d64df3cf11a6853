import torch
from .conv import MessagePassing  # noqa: F401
from .dense.linear import Linear  # noqa: F401
from .pool import global_mean_pool, global_max_pool, global_add_pool  # noqa: F401
from .sequential import Sequential  # noqa: F401
from . import inits, conv, dense, pool  # noqa: F401


class BatchNorm(torch.nn.Module):
    """gnn.BatchNorm: a wrapper whose only child is `.module = nn.BatchNorm1d` (key names
    `module.weight`, ... confirmed by the reference's HL_HGAT_Brain.pt checkpoint)."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = torch.nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def reset_parameters(self):
        self.module.reset_parameters()

    def forward(self, x):
        return self.module(x)
