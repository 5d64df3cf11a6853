#!/usr/bin/env python
"""bench.py -- headline benchmark of the HL-HGAT hot path on B200 (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one training pass (forward + backward + gradient all-reduce + Adam) of the reference's
ZINC model (HL_HGCNN_zinc_dense_int3_pyr, channels [2,2,2], filters [64,128,256], K=2, fp32) over one
synthetic ZINC-shaped mini-batch of 1024 graphs PER GPU (BASELINE.json configs[1]; weak scaling).
Prints ONE JSON line (rank 0).  `--workload peptides|cifar|tsp` runs the other BASELINE.json configs
(hlhgat_b200/workloads.py) through the same harness.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_OUT = sys.stdout

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

POOL = 4                      # distinct synthetic batches cycled through


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.proc = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.proc = subprocess.Popen(["nvidia-smi", "-i", uuid, f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    return world, rank, local


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle restatement of the reference on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_throughput(wl, steps, warmup, sample_batch):
    """The oracle restatement of the reference model of workload `wl` (same class name, same ctor arguments,
    same loss, Adam) on `sample_batch`-graph batches, all host cores."""
    from oracle import hodge_oracle as O
    from hlhgat_b200.workloads import focal_loss
    torch.manual_seed(0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = getattr(O, wl.model)(**wl.ctor).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    batches = [wl.make(sample_batch, 100 + i) for i in range(2)]

    def loss_of(b):
        if wl.model.endswith("attpool"):
            out = model(b, if_att=True)[0]
            y = b[0].y
            return torch.nn.functional.cross_entropy(out, y) if y.dtype == torch.int64 else focal_loss(out, y)
        if "TSP" in wl.model:
            return focal_loss(model(b)[0], b.y.view(-1, 1))
        return torch.nn.functional.l1_loss(model(b), b.y)

    def step(b):
        opt.zero_grad()
        loss = loss_of(b)
        loss.backward()
        opt.step()
        return loss.item()

    for i in range(warmup):
        step(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        step(batches[i % 2])
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps * 1e3, cores


def run_reference(args, world, rank):
    if rank != 0:
        return
    from hlhgat_b200.workloads import WORKLOADS
    wl = WORKLOADS[args.workload]
    sample = wl.cpu_sample
    gps, ms, cores = cpu_reference_throughput(wl, args.steps, args.warmup, sample)
    line = {"impl": "reference", "metric": wl.metric, "value": gps, "unit": "graphs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.label, "graphs_per_gpu": sample, "global_batch": sample,
                       "note": "CPU: pure-torch restatement of the reference (PyG not installable), not PyG itself"},
            "cpu_baseline": {"value": gps, "unit": "graphs/s", "cores": cores, "kind": "port",
                             "sample": (f"{args.steps} training steps on full {sample}-graph batches (the workload's own batch size), "
                                        if sample == wl.batch else
                                        f"{args.steps} training steps on {sample}-graph batches (a bounded sample of the "
                                        f"{wl.batch}-graph workload), ") + f"{cores} torch threads"},
            "e2e": {"value": gps, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# roofline leg: the polynomial SpMM kernel on an operator stack larger than L2
# ------------------------------------------------------------------------------------------------
ROOFLINE_CASES = {
    # workload: (block-diagonal replication, feature width): sized so inputs + outputs exceed the 126 MB L2
    "zinc": (16, 64), "zinc_default": (16, 64), "peptides": (40, 64), "cifar": (4, 64), "tsp": (4, 32),
}


def _time_launch(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in evs)


def spmm_roofline(dev, batch_dev, workload="zinc", factored=False, iters=20):
    """The polynomial SpMM launch of the step (HL_EPI_LAGUERRE_FIRST: T_1 = x - A x, the first SpMM of every conv) in two
    regimes: on the bench batch replicated `reps` times block-diagonally so that x and T_1 exceed the 126 MB L2 (the
    roofline figure), and on the bench batch itself (`in_step`: what a launch inside the timed step sees -- operator and
    features L2-resident).  CSR workloads: the fused L0+L1 launch.  `factored` (cifar / tsp): the edge operator is applied
    as diag(2/lambda) B1^T (B1 x) by hodge1_node_kernel + hodge1_edge_kernel -- THOSE launches are timed, against their
    own algorithmic bytes."""
    from hlhgat_b200 import functional as F_hl, _native as N
    from hlhgat_b200.simplex import CsrOperator, Hodge1Factor, incidence_for
    reps, width = ROOFLINE_CASES[workload]
    pk, kind = peaks()

    def build(rep):
        ops, xs, nnz_tot, rows_tot = [], [], 0, 0
        n0, e0 = batch_dev.x_t.shape[0], batch_dev.x_s.shape[0]
        for ei, ew, r in ((batch_dev.edge_index_t, batch_dev.edge_weight_t, n0), (batch_dev.edge_index_s, batch_dev.edge_weight_s, e0)):
            off = (torch.arange(rep, device=dev) * r).repeat_interleave(ei.shape[1])
            op = CsrOperator(ei.repeat(1, rep) + off, ew.repeat(rep), r * rep)
            op.fwd
            ops.append(op)
            xs.append(torch.randn(r * rep, width, device=dev))
            nnz_tot += ei.shape[1] * rep
            rows_tot += r * rep
        if factored:                                      # the edge operator in factored form: its own launches, its own bytes
            und = batch_dev.edge_index
            off = (torch.arange(rep, device=dev) * n0).repeat_interleave(und.shape[1])
            inc = incidence_for(und.repeat(1, rep) + off, n0 * rep)
            ops[1].factored = Hodge1Factor.from_operator(ops[1], inc)
            n, e = n0 * rep, e0 * rep
            node_pass = 4 * (n + 1) + 16 * e + 4 * width * (e + n)
            edge_pass = 12 * e + 4 * width * (n + e * (1 + 1))        # n_epi = 1: the own-row x of T_1 = x - A x
            return [ops[1]], [xs[1]], node_pass + edge_pass, e, nnz_tot
        alg = 8 * nnz_tot + 4 * (rows_tot + 2) + 4 * rows_tot * width * 2      # n_in = 1 (x), n_out = 1 (T_1)
        return ops, xs, alg, rows_tot, nnz_tot

    F_hl.enable_factored_hodge1(factored)
    ops, xs, alg_bytes, rows_tot, nnz_tot = build(reps)
    ms = _time_launch(lambda: F_hl.poly_basis_fwd(N.HL_LAGUERRE, 2, ops, xs, width), iters)
    ach = alg_bytes / (ms * 1e-3) / 1e9
    kernel = ("factored edge Laplacian: hodge1_node_kernel + hodge1_edge_kernel<LAGUERRE_FIRST> (the launches of the timed step)"
              if factored else "polynomial SpMM, fused L0+L1 launch (poly_spmm_staged_kernel / poly_spmm_kernel, chosen per operator)")
    out = {"bound": "hbm", "kernel": kernel,
           "achieved": ach, "peak": pk["hbm_gbs"], "peak_kind": f"{kind} (MEASURED_PEAKS.json hbm_gbs, burst copy)",
           "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
           "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": ms * 1e3,
           "workload": f"{workload}-shaped bench batch x{reps} block-diagonal, rows {rows_tot}, nnz(CSR) {nnz_tot}, F={width}; "
                       "inputs+outputs > L2"}
    del ops, xs
    ops1, xs1, alg1, rows1, _ = build(1)
    reps_g = 20
    launch1 = lambda: F_hl.poly_basis_fwd(N.HL_LAGUERRE, 2, ops1, xs1, width)      # noqa: E731
    for _ in range(3):
        launch1()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):            # 20 launches per replay: no host launch gaps, as inside the step graph
            for _ in range(reps_g):
                launch1()
    torch.cuda.synchronize()
    ms1 = _time_launch(graph.replay, 5) / reps_g
    out["in_step"] = {"achieved": alg1 / (ms1 * 1e-3) / 1e9, "unit": "GB/s", "us_per_launch": ms1 * 1e3,
                      "algorithmic_bytes_per_launch": alg1, "rows": rows1,
                      "note": "the same launch on the bench batch itself (20 launches per CUDA-graph replay, as inside the step graph): "
                              "operator and features are L2-resident, so this is an L2-bandwidth / launch-latency figure, not an HBM one"}
    t = ncu_traffic(workload, factored)
    if t is not None:
        out["traffic"], out["traffic_source"] = t
    return out


def ncu_traffic(workload, factored):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, from the committed `ncu --set full`
    capture of this round (the same operator stack; ncu cannot run inside the bench process)."""
    name = {("zinc", False): "r2_spmm_zinc_full.txt", ("zinc_default", False): "r2_spmm_zinc_full.txt"}.get((workload, factored))
    if name is None:
        return None
    try:
        tot = 0.0
        for ln in open(os.path.join(ROOT, "profiles", name)):
            if ln.startswith("dram__bytes_read.sum") or ln.startswith("dram__bytes_write.sum"):
                val, unit = ln.split("=")[1].split()[:2]
                tot += float(val) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit]
        return (tot, f"profiles/{name} (ncu --set full of tools/spmm_probe.py, same operator stack)") if tot else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def run_ours(args, world, rank, local):
    import hlhgat_b200
    from hlhgat_b200 import _native as N
    from hlhgat_b200.lib import Hodge_ST_Model as M
    from hlhgat_b200.parallel import FlatGradBucket, FlatAdam, broadcast_parameters
    from hlhgat_b200.synthetic import batch_to
    from hlhgat_b200.workloads import WORKLOADS
    from hlhgat_b200.training import Capacity, pad_batch, pad_levels, padded_nbytes, GraphedTrainStep, BatchPrefetcher

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    hlhgat_b200.build()
    wl = WORKLOADS[args.workload]
    batch_size = args.batch or wl.batch
    factored = args.factored_l1 == "on" or (args.factored_l1 == "auto" and wl.long_rows)
    from hlhgat_b200 import functional as F_hl
    F_hl.enable_factored_hodge1(factored)
    hlhgat_b200.enable_lanes(args.lanes == "on")
    hlhgat_b200.enable_project_then_transfer(args.project_first == "on")
    from hlhgat_b200.dense_stack import enable_dense_stack
    enable_dense_stack(args.dense_stack == "on")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = getattr(M, wl.model)(**wl.ctor).to(dev).train()
    broadcast_parameters(model)
    bucket = FlatGradBucket(model.parameters())
    if args.optimizer == "flat":      # torch.optim.Adam's update rule as one streaming kernel over flat buffers (hl_adam_flat)
        opt = FlatAdam(bucket, lr=1e-3, weight_decay=1e-3)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3, fused=True, capturable=True)

    raw = [wl.make(batch_size, 1000 * rank + i) for i in range(args.pool)]
    levels = [[b[l] for b in raw] for l in range(wl.levels)] if wl.levels > 1 else [raw]
    caps = [Capacity.covering(lv) for lv in levels]
    if world > 1:                                        # same capacity on every rank (same graph shapes)
        t = torch.tensor([[c.nodes, c.edges, c.nnz_t, c.nnz_s] for c in caps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for c, row in zip(caps, t.tolist()):
            c.nodes, c.edges, c.nnz_t, c.nnz_s = (int(v) for v in row)
    # pinned host buffers in the reference's batch format, padded to the fixed capacity
    if wl.levels > 1:
        host = [pad_levels(b, caps, pin=True, deg_eps=wl.deg_eps) for b in raw]
    else:
        host = [pad_batch(b, caps[0], pin=True, deg_eps=wl.deg_eps) for b in raw]
    h2d_bytes = padded_nbytes(host[0])
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    stepper = GraphedTrainStep(model, wl.loss, opt, bucket, host[0], dev, warmup=3, loss_fn=True)
    resident = []
    for b in host:
        stepper.batch.load(b)
        resident.append(stepper.batch.clone_resident())
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def step_resident(i):                                # inputs already in HBM: D2D into the graph's static buffers
        stepper.batch.load(resident[i % args.pool])
        stepper.step()

    prefetch = BatchPrefetcher(stepper.batch, host[0], dev)
    loss_slots = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_events = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"primed": False, "losses": []}

    def step_e2e(i):                                     # host buffers in, loss out, every step
        # Every step's batch comes from pinned host memory; its copy is issued one step ahead on a copy stream, so it
        # overlaps the previous step's kernels (the first step of a run pays for its own copy).  Every step's loss is
        # read back to the host; the host consumes it one step later, so the device never idles on the read.
        if not state["primed"]:
            prefetch.prefetch(host[i % args.pool], i % 2)
            state["primed"] = True
        prefetch.swap_in(i % 2)
        prefetch.prefetch(host[(i + 1) % args.pool], (i + 1) % 2)
        loss = stepper.step()
        loss_slots[i % 2].copy_(loss, non_blocking=True)
        loss_events[i % 2].record()
        if i > 0:
            loss_events[(i - 1) % 2].synchronize()
            state["losses"].append(float(loss_slots[(i - 1) % 2]))

    def drain_e2e(steps):
        loss_events[(steps - 1) % 2].synchronize()
        state["losses"].append(float(loss_slots[(steps - 1) % 2]))
        loss_host.copy_(loss_slots[(steps - 1) % 2])

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    for i in range(args.warmup):
        step_e2e(i)
    drain_e2e(args.warmup)
    state["primed"] = False
    prefetch.stream.synchronize()
    state["losses"].clear()

    def e2e_run(i):
        step_e2e(i)
        if i == args.steps - 1:
            drain_e2e(args.steps)                        # the last loss is on the host before the closing timestamp
    ms_e2e = timed(e2e_run, args.steps)
    assert len(state["losses"]) == args.steps
    launches = stepper.launches_per_step * args.steps
    final_loss = float(loss_host)

    value = batch_size * world * args.steps / (ms * 1e-3)
    e2e = batch_size * world * args.steps / (ms_e2e * 1e-3)
    if rank != 0:
        return
    cap_txt = " + ".join(f"{c.nodes} nodes / {c.edges} edges" for c in caps)
    line = {"metric": wl.metric, "value": value, "unit": "graphs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.label + (f"_batch{batch_size}" if batch_size != wl.batch else ""), "graphs_per_gpu": batch_size, "global_batch": batch_size * world,
                       "parallelism": f"dp{world}", "batch_pool": args.pool,
                       "l2": "no explicit flush: per-step working set (activations saved for backward, ~1 GB or more) exceeds the "
                             "126 MB L2 and consecutive steps use different batches",
                       "execution": "whole step (CSR bucketing + forward + backward) replayed as one CUDA graph on batches padded "
                                    f"to a fixed capacity ({cap_txt}, ~2% ghost rows), then all-reduce + fused Adam graph"
                                    + ("; node chain and edge chain of every layer on two streams = two parallel branches of the graph"
                                       if args.lanes == "on" else ""),
                       "dense_connections": ("preallocated [rows, width] buffers: BatchNorm writes each block in place, each block transferred once, "
                                             "data gradients accumulated by the GEMM epilogue (forward bit-identical to cat + re-transfer)"
                                             if args.dense_stack == "on" else "torch.cat per layer + whole concat re-transferred (the reference's scheme)"),
                       "transfer": ("project-then-transfer in NodeEdgeInt: (1/D)|B1| (x W_a^T) instead of ((1/D)|B1| x) W_a^T"
                                    if args.project_first == "on" else "transfer-then-project (the reference's order)"),
                       "edge_operator": ("L1 applied in factored form diag(2/lambda) B1^T B1 (opt-in, fp32-rounding-equal to the CSR path)"
                                         if factored else "L1 applied from its CSR (bit-exact summation order of the reference)"),
                       "optimizer": ("Adam(lr 1e-3, weight_decay 1e-3) as one hl_adam_flat launch over flat parameter / gradient / moment "
                                     "buffers, 1/world_size folded in" if args.optimizer == "flat" else "torch.optim.Adam(fused=True, capturable=True)"),
                       "gemm": "dense Theta/MLP transforms + data/weight gradients: hand-written tcgen05 3xTF32 kernels (fp32-accurate); "
                               "cuBLAS fp32 only for shapes with N % 16 != 0 or unaligned rows (first-layer inputs)"},
            "e2e": {"value": e2e, "unit": "graphs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "pipeline": "H2D of batch i+1 on a copy stream during step i (two pinned->device staging slots, D2D into the "
                                "graph's static buffers); the loss of every step is copied to pinned host memory and read by the "
                                "host one step later"},
            "gpu_launches": int(launches), "gpu_launches_note": "libhlhgat kernels inside the replayed graph x steps (cuBLAS/ATen launches not counted)",
            "final_loss": final_loss, "clocks": clocks}
    try:
        b0 = raw[0][0] if wl.levels > 1 else raw[0]
        line["roofline"] = spmm_roofline(dev, batch_to(b0, dev), args.workload, factored)
    except Exception as exc:  # pragma: no cover
        line["roofline"] = {"error": repr(exc)}
    try:
        line["roofline_dense"] = dense_roofline(dev, caps[0].nodes)
    except Exception as exc:  # pragma: no cover
        line["roofline_dense"] = {"error": repr(exc)}
    if world == 1 and not args.no_cpu_baseline:
        gps, ms_cpu, cores = cpu_reference_throughput(wl, 3, 1, wl.cpu_sample)
        line["cpu_baseline"] = {"value": gps, "unit": "graphs/s", "cores": cores, "kind": "port",
                                "sample": f"3 training steps (1 warm-up) on {wl.cpu_sample}-graph batches "
                                          + ("(the workload's full batch size)" if wl.cpu_sample == wl.batch else f"(bounded sample of the {wl.batch}-graph batch)")
                                          + f" of the same shape, same model, {cores} torch threads; pure-torch restatement of the reference, not PyG"}
    print(json.dumps(line), file=_OUT, flush=True)


def dense_roofline(dev, rows, iters=20):
    """Second roofline leg (the dense Theta / MLP transform is the largest share of the step's kernel time): the
    tcgen05 3xTF32 GEMM on the largest dense shape of the ZINC stack, [rows,1408] x [1408,256]^T (first MLP Linear of
    the last NodeEdgeInt; A = 136 MB > L2).  fp32-equivalent FLOPs 2*M*N*K over the CUDA-event time; every
    product costs three tf32 MMAs and tf32 runs at half the bf16 rate, so the tensor-pipe ceiling for this arithmetic
    is bf16_peak / 6."""
    from hlhgat_b200 import _native as N
    L = N.lib()
    M, Nn, K = int(rows), 256, 1408
    a = torch.randn(M, K, device=dev)
    w = torch.randn(Nn, K, device=dev) * 0.05
    hi, lo, c = torch.empty_like(w), torch.empty_like(w), torch.empty(M, Nn, device=dev)
    N.check(L.hl_tf32_split(w.data_ptr(), K, Nn, K, 0, hi.data_ptr(), lo.data_ptr(), K, N.stream_ptr()), "hl_tf32_split")

    def launch():
        rc = L.hl_gemm_tf32x3(a.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, M, Nn, K, None, c.data_ptr(), Nn, 0, N.stream_ptr())
        if rc != 0:
            raise RuntimeError(f"hl_gemm_tf32x3 returned {rc}")
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for e0, e1 in evs:
        e0.record()
        launch()
        e1.record()
    torch.cuda.synchronize()
    ms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in evs)
    pk, kind = peaks()
    ach = 2.0 * M * Nn * K / (ms * 1e-3) / 1e12
    peak = pk["bf16_tflops"] / 6.0
    err = float((c[:256].double() - a[:256].double() @ w.double().t()).abs().max() / (a[:256].double() @ w.double().t()).abs().max())
    return {"bound": "tensor", "kernel": "gemm_tf32x3_kernel<0> (tcgen05.mma kind::tf32, 3 MMAs per product, fp32 TMEM accumulator)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_kind": f"{kind} dense bf16 {pk['bf16_tflops']:.0f} TFLOP/s / 2 (tf32 rate) / 3 (MMAs per fp32-accurate product)",
            "us_per_launch": ms * 1e3, "shape": [M, Nn, K], "max_rel_err_vs_fp64": err, "traffic": None}


def _claim_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on stdout.
    Keep a private handle to the real stdout and point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="zinc", choices=["zinc", "zinc_default", "peptides", "cifar", "tsp"],
                    help="BASELINE.json config: zinc = configs[1] (the headline metric, default); the others are the "
                         "peptides-func / CIFAR10-superpixel / TSP-shaped configs")
    ap.add_argument("--batch", type=int, default=0,
                    help="graphs per GPU (default: the workload's own: 1024 zinc, 64 peptides (script default; SURVEY also lists "
                         "256), 256 cifar, 32 tsp)")
    ap.add_argument("--pool", type=int, default=POOL, help="distinct synthetic batches cycled through")
    ap.add_argument("--lanes", default="on", choices=["on", "off"],
                    help="issue the node chain and the edge chain of every layer on two CUDA streams (bit-identical results)")
    ap.add_argument("--project-first", default="off", choices=["on", "off"],
                    help="NodeEdgeInt applies W_a before the node<->edge transfer when the layer is narrower than the "
                         "dense-connection buffer (transfers move f instead of d columns; fp32-rounding-equal)")
    ap.add_argument("--optimizer", default="flat", choices=["flat", "torch"],
                    help="flat: Adam (torch.optim.Adam semantics, lr 1e-3, weight decay 1e-3 as in main_zinc...py:213) as one "
                         "hand-written kernel over flat parameter / gradient / moment buffers; torch: torch.optim.Adam(fused, capturable)")
    ap.add_argument("--dense-stack", default="on", choices=["on", "off"],
                    help="dense connections in preallocated buffers (no torch.cat, every block transferred to the other simplex "
                         "order once, data gradients accumulated in the GEMM epilogue); off = the reference's cat + full re-transfer")
    ap.add_argument("--factored-l1", default="auto", choices=["auto", "on", "off"],
                    help="apply the edge Laplacian as diag(2/lambda) B1^T B1 instead of its CSR; auto = only for the long-row "
                         "workloads (cifar, tsp); the ZINC headline always uses the CSR SpMM")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world, rank, local = dist_setup(args.gpus)
    try:
        if args.impl == "reference":
            run_reference(args, world, rank)
        else:
            run_ours(args, world, rank, local)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
