import torch


def global_mean_pool(x, batch, size=None):
    if batch is None:
        return x.mean(dim=0, keepdim=True)
    b = int(batch.max()) + 1 if size is None else size
    out = torch.zeros((b,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    cnt = torch.zeros((b,), dtype=x.dtype, device=x.device).index_add_(
        0, batch, torch.ones(batch.size(0), dtype=x.dtype, device=x.device))
    return out / cnt.clamp(min=1).view((-1,) + (1,) * (x.dim() - 1))


def global_max_pool(x, batch, size=None):
    b = int(batch.max()) + 1 if size is None else size
    idx = batch.view((-1,) + (1,) * (x.dim() - 1)).expand_as(x)
    out = torch.full((b,) + tuple(x.shape[1:]), float("-inf"), dtype=x.dtype, device=x.device)
    return out.scatter_reduce_(0, idx, x, "amax", include_self=True)


def global_add_pool(x, batch, size=None):
    b = int(batch.max()) + 1 if size is None else size
    return torch.zeros((b,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device).index_add_(0, batch, x)


def graclus(*a, **k):  # imported by the reference but never called
    raise NotImplementedError("graclus is not on the hot path (SURVEY.md K12)")


def max_pool(*a, **k):
    raise NotImplementedError
