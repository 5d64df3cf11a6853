"""Every dense launch (forward / data-gradient GEMM, weight gradient) of ONE training step of a bench workload, with
its shape recorded from the real step, then timed shape by shape inside a CUDA graph (20 launches per replay) and set
against its own floor: max(flops / the 3xTF32 ceiling, bytes / measured HBM peak).  The step-weighted fraction at the end
is sum(floor) / sum(time) over all launches of the step.

  python tools/dense_shapes_probe.py [workload] > gpurun_out/dense_shapes_<workload>.txt
"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hlhgat_b200  # noqa: E402
from hlhgat_b200 import functional as F  # noqa: E402
from hlhgat_b200 import dense_stack as DS  # noqa: E402
from hlhgat_b200.lib import Hodge_ST_Model as M  # noqa: E402
from hlhgat_b200.simplex import clear_caches  # noqa: E402
from hlhgat_b200.training import Capacity, StaticBatch, pad_batch, pad_levels  # noqa: E402
from hlhgat_b200.workloads import WORKLOADS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "zinc"
wl = WORKLOADS[name]
dev = torch.device("cuda:0")
try:
    pk = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    HBM, TF = pk["hbm_gbs"], pk["bf16_tflops"] / 6.0
except Exception:
    HBM, TF = 6467.7, 281.3

shapes = collections.Counter()
orig = dict(dense=F.dense, dense2=F.dense2, wgrad=F.wgrad, wgrad2=F.wgrad2)


def rec_dense(a, w, bias=None, out=None, accumulate=False, transpose_w=False, bn=None):
    n_out = w.shape[1] if transpose_w else w.shape[0]
    shapes[("gemm", a.shape[0], n_out, a.shape[1], 0, bool(accumulate))] += 1
    return orig["dense"](a, w, bias, out, accumulate, transpose_w, bn)


def rec_dense2(a1, w1, a2, w2, bias=None, out=None, accumulate=False, bn=None):
    shapes[("gemm", a1.shape[0], w1.shape[0], a1.shape[1], a2.shape[1], bool(accumulate))] += 1
    return orig["dense2"](a1, w1, a2, w2, bias, out, accumulate, bn)


def rec_wgrad(g, x, out=None, accumulate=False, bias_out=None, bias_accumulate=False):
    shapes[("wgrad", g.shape[0], g.shape[1], x.shape[1], 0, False)] += 1
    return orig["wgrad"](g, x, out, accumulate, bias_out, bias_accumulate)


def rec_wgrad2(g, x1, x2, out1, out2, accumulate=False, bias_out=None, bias_accumulate=False):
    shapes[("wgrad", g.shape[0], g.shape[1], x1.shape[1], x2.shape[1], False)] += 1
    return orig["wgrad2"](g, x1, x2, out1, out2, accumulate, bias_out, bias_accumulate)


torch.manual_seed(0)
F.enable_factored_hodge1(wl.long_rows)
model = getattr(M, wl.model)(**wl.ctor).to(dev).train()
raw = wl.make(wl.batch, 0)
if wl.levels > 1:
    host = pad_levels(raw, [Capacity.covering([raw[l]]) for l in range(wl.levels)], deg_eps=wl.deg_eps)
else:
    host = pad_batch(raw, Capacity.covering([raw]), deg_eps=wl.deg_eps)
batch = StaticBatch(host, dev)
clear_caches()
wl.loss(model, batch).backward()               # warm-up (weight splits, caches)
F.dense, F.dense2, F.wgrad, F.wgrad2 = rec_dense, rec_dense2, rec_wgrad, rec_wgrad2
clear_caches()
wl.loss(model, batch).backward()
F.dense, F.dense2, F.wgrad, F.wgrad2 = orig["dense"], orig["dense2"], orig["wgrad"], orig["wgrad2"]
torch.cuda.synchronize()
del model, batch
torch.cuda.empty_cache()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


print(f"workload {name}: dense launches of one training step ({sum(shapes.values())} launches, {len(shapes)} distinct shapes); "
      f"floors against {TF:.1f} TFLOP/s (bf16 / 6) and {HBM:.0f} GB/s")
print(f"{'kind':6s} {'rows':>8s} {'N/fo':>5s} {'K/fi':>5s} {'K2':>5s} acc  x   us/launch  TFLOP/s  flop-floor  hbm-floor  frac")
tot_t = tot_floor = tot_fl = 0.0
by_kind = collections.defaultdict(lambda: [0.0, 0.0])
for (kind, R, n, k, k2, acc), cnt in sorted(shapes.items(), key=lambda kv: (kv[0][0], -kv[0][1] * kv[0][2] * (kv[0][3] + kv[0][4]))):
    if R < 1:
        continue
    if kind == "gemm":
        a1 = torch.randn(R, k, device=dev)
        w1 = torch.randn(n, k, device=dev) * 0.1
        out = torch.empty(R, n, device=dev)
        if k2:
            a2 = torch.randn(R, k2, device=dev)
            w2 = torch.randn(n, k2, device=dev) * 0.1
            fn = lambda: F.dense2(a1, w1, a2, w2, None, out=out, accumulate=acc)      # noqa: E731
        else:
            fn = lambda: F.dense(a1, w1, None, out=out, accumulate=acc)               # noqa: E731
        fl = 2.0 * R * n * (k + k2)
        by = 4.0 * (R * (k + k2) + R * n * (2 if acc else 1))
    else:
        g = torch.randn(R, n, device=dev)
        x1 = torch.randn(R, k, device=dev)
        o1 = torch.empty(n, k, device=dev)
        if k2:
            x2 = torch.randn(R, k2, device=dev)
            o2 = torch.empty(n, k2, device=dev)
            fn = lambda: F.wgrad2(g, x1, x2, o1, o2)                                  # noqa: E731
        else:
            fn = lambda: F.wgrad(g, x1, o1)                                           # noqa: E731
        fl = 2.0 * R * n * (k + k2)
        by = 4.0 * (R * n + R * (k + k2))
    us = timed(fn)
    f_fl, f_hbm = fl / TF / 1e6, by / HBM / 1e3
    floor = max(f_fl, f_hbm)
    print(f"{kind:6s} {R:8d} {n:5d} {k:5d} {k2:5d} {int(acc):3d} {cnt:3d} {us:9.1f} {fl / us / 1e6:8.1f} {f_fl:10.1f} {f_hbm:10.1f} {floor / us:6.2f}",
          flush=True)
    tot_t += cnt * us
    tot_floor += cnt * floor
    tot_fl += cnt * fl
    by_kind[kind][0] += cnt * us
    by_kind[kind][1] += cnt * floor
    a1 = w1 = out = a2 = w2 = g = x1 = x2 = o1 = o2 = None
for kind, (t, fl) in by_kind.items():
    print(f"{kind}: {t / 1e3:.2f} ms per step alone, floor {fl / 1e3:.2f} ms, fraction {fl / t:.2f}")
print(f"all dense launches: {tot_t / 1e3:.2f} ms per step when run alone, floor {tot_floor / 1e3:.2f} ms, step-weighted fraction {tot_floor / tot_t:.2f}, "
      f"{tot_fl / tot_t / 1e6:.1f} TFLOP/s")
