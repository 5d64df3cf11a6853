"""Whole-step CUDA-graph execution for the launch-bound small-graph regime (SURVEY.md section 7,
"hard parts": at ZINC batch 1024 one step is ~900 kernels of a few microseconds each).

Mini-batches differ in their node / edge / nnz totals, so they are padded on the host to a fixed
*capacity* with inert "ghost" rows -- zero features, zero operator weights, their own ghost graph --
and the true row counts travel as device scalars.  BatchNorm is the only place where row counts
matter: the kernels exclude rows beyond `*nvalid` from the statistics and write exact zeros there
(forward and backward), so ghost rows contribute nothing to any reduction or weight gradient and
results equal the unpadded batch.  With every shape static, forward + backward (+ the CSR bucketing
of the freshly copied COO) replay as ONE cudaGraphLaunch; the gradient all-reduce and the fused Adam
step follow.
"""
from types import SimpleNamespace

import torch

from . import lanes
from .functional import accumulate_into_grads, WeightSplitPlan, WgradReducePlan
from .simplex import clear_caches

_PAD_KEYS = ("x_t", "x_s", "y", "edge_index", "edge_index_t", "edge_index_s", "edge_weight_t", "edge_weight_s",
             "num_node1", "num_edge1", "D", "n_valid_nodes", "n_valid_edges", "n_valid_graphs")


class Capacity(SimpleNamespace):
    """Fixed sizes of a padded batch: nodes, edges, nnz_t, nnz_s (each strictly larger than any real
    batch so that a ghost row exists) and graphs (real graphs; one ghost graph is appended)."""

    @classmethod
    def covering(cls, batches, slack=0.02):
        def cap(values):
            m = max(values)
            return int(m + max(8, int(m * slack)) + 7) // 8 * 8
        return cls(nodes=cap([b.x_t.shape[0] for b in batches]), edges=cap([b.x_s.shape[0] for b in batches]),
                   nnz_t=cap([b.edge_index_t.shape[1] for b in batches]),
                   nnz_s=cap([b.edge_index_s.shape[1] for b in batches]),
                   graphs=max(b.num_graphs for b in batches))


def pad_batch(b, cap, pin=False, deg_eps=0.0):
    """Host-side padding of a collated batch to `cap` (data preparation, like the collate itself).
    `deg_eps`: the constant the model adds to the node degree (0 for the ZINC model, 1e-6 elsewhere);
    a target `y` with one row per edge (TSP) is padded like the edge features."""
    n, e = b.x_t.shape[0], b.x_s.shape[0]
    g = b.num_graphs
    assert n < cap.nodes and e < cap.edges and g == cap.graphs
    assert b.edge_index_t.shape[1] <= cap.nnz_t and b.edge_index_s.shape[1] <= cap.nnz_s

    def rows(x, total):
        out = torch.zeros((total,) + tuple(x.shape[1:]), dtype=x.dtype)
        out[: x.shape[0]] = x
        return out

    def cols(ei, total, first_ghost, n_ghost):
        """pad a [2,nnz] index list with self-pairs spread round-robin over the ghost rows (a single
        ghost row would serialise thousands of padding entries in one lane group)"""
        out = torch.empty((2, total), dtype=ei.dtype)
        out[:, : ei.shape[1]] = ei
        fill = first_ghost + torch.arange(total - ei.shape[1]) % n_ghost
        out[:, ei.shape[1]:] = fill
        return out

    p = SimpleNamespace(num_graphs=g)
    p.x_t, p.x_s = rows(b.x_t, cap.nodes), rows(b.x_s, cap.edges)
    p.y = rows(b.y, cap.edges) if (b.y.dim() and b.y.shape[0] == e and e != g) else b.y.clone()
    p.edge_index = cols(b.edge_index, cap.edges, n, cap.nodes - n)          # ghost edges: self-pairs on ghost nodes
    p.edge_index_t = cols(b.edge_index_t, cap.nnz_t, n, cap.nodes - n)      # ghost entries: (g, g) with weight 0
    p.edge_index_s = cols(b.edge_index_s, cap.nnz_s, e, cap.edges - e)
    p.edge_weight_t, p.edge_weight_s = rows(b.edge_weight_t, cap.nnz_t), rows(b.edge_weight_s, cap.nnz_s)
    p.num_node1 = torch.cat([b.num_node1, torch.tensor([cap.nodes - n])])
    p.num_edge1 = torch.cat([b.num_edge1, torch.tensor([cap.edges - e])])
    deg = torch.zeros(cap.nodes).index_add_(0, b.edge_index.reshape(-1), torch.ones(2 * e))
    deg = deg + deg_eps
    deg[n:] = 1.0                                                # ghost nodes: finite 1/D
    p.D = deg
    p.n_valid_nodes = torch.tensor([n], dtype=torch.int32)
    p.n_valid_edges = torch.tensor([e], dtype=torch.int32)
    p.n_valid_graphs = torch.tensor([g], dtype=torch.int32)
    if pin:
        for k in _PAD_KEYS:
            setattr(p, k, getattr(p, k).pin_memory())
    return p


def padded_nbytes(p):
    if isinstance(p, (list, tuple)):
        return sum(padded_nbytes(q) for q in p)
    return sum(getattr(p, k).numel() * getattr(p, k).element_size() for k in _PAD_KEYS)


def pad_levels(datas, caps, pin=False, deg_eps=1e-6):
    """Two-level batches of the attention-pooling models: every level is padded on its own.  Ghost
    rows of level 0 carry cluster id 0 inside the ghost graph, i.e. they pool into the first ghost row
    of level 1 (all-zero features either way)."""
    return [pad_batch(d, c, pin, deg_eps) for d, c in zip(datas, caps)]


class StaticBatch:
    """Device buffers of one capacity; `load` overwrites them in place (H2D or D2D copies).  A list of
    padded levels gives a list-like StaticBatch (`batch[0]`, `batch[1]`)."""

    def __init__(self, proto, device):
        self.levels = None
        if isinstance(proto, (list, tuple)):
            self.levels = [StaticBatch(p, device) for p in proto]
            return
        self.num_graphs = proto.num_graphs
        for k in _PAD_KEYS:
            setattr(self, k, torch.empty_like(getattr(proto, k), device=device))
        self.load(proto)

    def __getitem__(self, i):
        return self.levels[i]

    def __len__(self):
        return len(self.levels)

    def load(self, p, non_blocking=True):
        if self.levels is not None:
            for lv, q in zip(self.levels, p):
                lv.load(q, non_blocking)
            return
        for k in _PAD_KEYS:
            getattr(self, k).copy_(getattr(p, k), non_blocking=non_blocking)

    def clone_resident(self):
        if self.levels is not None:
            return [lv.clone_resident() for lv in self.levels]
        out = SimpleNamespace(num_graphs=self.num_graphs)
        for k in _PAD_KEYS:
            setattr(out, k, getattr(self, k).clone())
        return out


class BatchPrefetcher:
    """Host -> device copies off the compute stream: while step i runs, batch i + 1 travels from pinned host memory into
    one of two staging `StaticBatch`es on a copy stream; `swap_in` then moves it into the buffers the captured step reads
    with device-to-device copies (microseconds), ordered by events in both directions."""

    def __init__(self, target, proto, device):
        self.target, self.device = target, device
        self.staging = [StaticBatch(proto, device), StaticBatch(proto, device)]
        self.stream = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]        # H2D into slot done
        self.free = [torch.cuda.Event(), torch.cuda.Event()]         # slot read out by the compute stream
        cur = torch.cuda.current_stream(device)
        for ev in self.free:
            ev.record(cur)

    def prefetch(self, host_batch, slot):
        self.stream.wait_event(self.free[slot])
        with torch.cuda.stream(self.stream):
            self.staging[slot].load(host_batch, non_blocking=True)
            self.ready[slot].record(self.stream)

    def swap_in(self, slot):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[slot])
        self.target.load(self.staging[slot])
        self.free[slot].record(cur)


class GraphedTrainStep:
    """forward + loss + backward captured once; replayed per step on the batch currently loaded in
    `self.batch`.  `optimizer` must be capturable (e.g. Adam(fused=True, capturable=True)).
    `criterion` is either a loss module applied to `(model(batch)[:num_graphs], batch.y)` (graph-level
    targets) or, with `loss_fn=True`, a callable `criterion(model, batch) -> scalar loss`."""

    def __init__(self, model, criterion, optimizer, bucket, proto_batch, device, warmup=3, loss_fn=False, single_graph=None):
        """single_graph (default: env HL_STEP_GRAPH != "split"): the gradient all-reduce (NCCL, capturable) and the optimizer
        step are captured into the SAME graph as forward + backward -- one cudaGraphLaunch per training step, no host
        round trips between backward, all-reduce and Adam.  False: three launches (graph, eager all-reduce, graph)."""
        import os
        self.model, self.criterion, self.optimizer, self.bucket = model, criterion, optimizer, bucket
        self.loss_fn = loss_fn
        self.device = device
        self.single_graph = (os.environ.get("HL_STEP_GRAPH", "single") != "split") if single_graph is None else bool(single_graph)
        self.batch = StaticBatch(proto_batch, device)
        self.loss = torch.zeros((), device=device)
        self.split_plan = WeightSplitPlan() if warmup >= 2 else None     # recorded by the first warm-up step
        self.reduce_plan = WgradReducePlan()
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup):                               # eager warm-up on a side stream
                self._fwd_bwd()
                self._reduce_and_update()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        clear_caches()
        from . import _native as N
        before = N.lib().hl_launch_count()
        self.graph_fb = torch.cuda.CUDAGraph()
        self.graph_opt = None
        with torch.cuda.graph(self.graph_fb):
            self._fwd_bwd()
            self.launches_per_step = int(N.lib().hl_launch_count() - before)   # libhlhgat kernels in the graph
            if self.single_graph:
                self._reduce_and_update()
        if not self.single_graph:
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph_fb.pool()):
                self.optimizer.step()
        clear_caches()

    def _reduce_and_update(self):
        from .parallel import FlatAdam
        if isinstance(self.optimizer, FlatAdam) and self.single_graph:
            self.bucket.all_reduce_sum()          # the 1 / world_size rides on the Adam kernel's gradient scale
        else:
            self.bucket.all_reduce_mean()
        self.optimizer.step()

    def _fwd_bwd(self):
        if self.split_plan is None:
            return self._fwd_bwd_body()
        with self.split_plan:              # all weight hi / lo splits in one launch on a side branch
            return self._fwd_bwd_body()

    def _fwd_bwd_body(self):
        clear_caches()                     # the static COO buffers change content between replays
        self.bucket.zero()
        if self.loss_fn:
            loss = self.criterion(self.model, self.batch)
        else:
            g = self.batch.num_graphs
            pred = self.model(self.batch, device=self.device)
            loss = self.criterion(pred[:g], self.batch.y)
        with self.reduce_plan:                     # ... their split reduces batched into one launch after the streams joined
            with accumulate_into_grads():          # weight gradients land in the flat bucket directly
                loss.backward()
        lanes.join(self.device)                     # edge-lane weight gradients land in the bucket before the all-reduce
        self.loss.copy_(loss.detach())

    def step(self):
        self.graph_fb.replay()
        if self.graph_opt is not None:
            self.bucket.all_reduce_mean()
            self.graph_opt.replay()
        return self.loss
