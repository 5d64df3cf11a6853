"""Data parallelism over graph mini-batches: one process per GPU,
replicated parameters, graphs sharded across ranks, ONE exchange per step -- an all-reduce of a
flat fp32 gradient bucket (2.5-3.9M parameters = 10-16 MB) over NCCL / NVLink."""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """Makes every parameter's .grad a view into one contiguous buffer so the whole gradient is
    exchanged with a single collective (latency-bound at this size: one launch, no bucketing)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    pending_scale = 1.0          # set by all_reduce_sum(): the 1 / world_size a FlatAdam step folds into its update

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))

    def all_reduce_sum(self, group=None):
        """SUM only; the division by the world size is left to the optimizer (`FlatAdam.step` reads `pending_scale`)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.pending_scale = 1.0 / dist.get_world_size(group)


class FlatAdam:
    """torch.optim.Adam semantics (L2 weight decay, bias correction) as ONE streaming kernel over flat buffers
    (`hl_adam_flat`): the parameters are moved into one contiguous buffer (every `p.data` becomes a view of it -- call
    this BEFORE the first forward pass, the weight-split plan keys on data pointers), the gradients are the views of
    `bucket` (a FlatGradBucket over the same parameters, same order), the moments are flat too.  The data-parallel mean is
    folded in: `step()` scales the (summed) gradients by 1 / world_size, so the bucket only has to all-reduce with SUM.
    Capturable: the step counter lives on the device."""

    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.bucket, self.lr, self.betas, self.eps, self.weight_decay = bucket, lr, betas, eps, weight_decay
        ps = bucket.params
        flat = torch.empty_like(bucket.flat)
        off = 0
        with torch.no_grad():
            for p in ps:
                n = p.numel()
                flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat[off:off + n].view_as(p)
                off += n
        self.flat_params = flat
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.state = torch.zeros(3, dtype=torch.float32, device=flat.device)

    def step(self, grad_scale=None):
        from . import _native as N
        if grad_scale is None:
            grad_scale = self.bucket.pending_scale
            self.bucket.pending_scale = 1.0
        N.check(N.lib().hl_adam_flat(self.flat_params.data_ptr(), self.bucket.flat.data_ptr(), self.exp_avg.data_ptr(),
                                     self.exp_avg_sq.data_ptr(), self.flat_params.numel(), self.state.data_ptr(), self.lr,
                                     self.betas[0], self.betas[1], self.eps, self.weight_decay, float(grad_scale), N.stream_ptr()),
                "hl_adam_flat")

    def zero_grad(self, set_to_none=False):
        self.bucket.zero()


def shard_graphs(costs, world_size):
    """Greedy longest-processing-time split of graph indices across ranks, balanced by cost
    (e.g. nnz(L1) per graph -- TSP/CIFAR graphs vary, ZINC barely).  Returns a list of index lists."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0] * world_size
    parts = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: loads[k])
        parts[r].append(i)
        loads[r] += costs[i]
    return [sorted(p) for p in parts]


def broadcast_parameters(module, src=0, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src, group=group)


# ---------------------------------------------------------------------------------------------
# BatchNorm statistics over all ranks (caveat 1).  Plain data parallelism normalises with
# per-rank statistics (what DDP does; the default here); the reference trains the whole batch on one GPU, so a
# run that must match a single-GPU run of the GLOBAL batch needs these two exchanges per BatchNorm layer.
# ---------------------------------------------------------------------------------------------
_SYNC_BN = {"group": None, "enabled": False}


def enable_sync_batchnorm(flag=True, group=None):
    """Opt in: every training-mode BatchNorm of the module layer (lib/Hodge_Cheb_Conv._bn_relu) uses statistics
    over all ranks of `group` (default: the world)."""
    _SYNC_BN["enabled"], _SYNC_BN["group"] = bool(flag), group


def sync_batchnorm_group():
    """The process group to synchronise BatchNorm over, or False when synchronisation is off / pointless."""
    if not _SYNC_BN["enabled"] or not (dist.is_available() and dist.is_initialized()):
        return False
    if dist.get_world_size(_SYNC_BN["group"]) < 2:
        return False
    return _SYNC_BN["group"]


def combine_bn_stats(local_stats, local_count, group=None):
    """Global (mean | biased variance) [2F] and row count from every rank's own: all-gather of [count, mean, var]
    (2F + 1 floats per rank), merged in fp64 -- mean = sum n_r m_r / N, var = sum n_r (v_r + (m_r - mean)^2) / N
    (no E[x^2] - mean^2 cancellation).  `local_count`: 0-dim / 1-element tensor on the same device."""
    f = local_stats.numel() // 2
    mine = torch.cat([local_count.reshape(1).to(local_stats.dtype), local_stats.reshape(-1)])
    world = dist.get_world_size(group)
    allr = torch.empty(world * (2 * f + 1), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(allr, mine.contiguous(), group=group)
    allr = allr.view(world, 2 * f + 1).double()
    n, m, v = allr[:, :1], allr[:, 1:f + 1], allr[:, f + 1:]
    total = n.sum()
    denom = total.clamp(min=1.0)
    mean = (n * m).sum(0) / denom
    var = (n * (v + (m - mean) ** 2)).sum(0) / denom
    return torch.cat([mean, var]).to(local_stats.dtype), total.to(local_stats.dtype)


def reduce_bn_sums(local_sums, group=None):
    """Column sums of dz and dz * xhat over all ranks (the per-rank ones stay this rank's dbeta / dgamma)."""
    out = local_sums.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
