"""Timeline of ONE replay of the whole-step CUDA graph (the bench configuration): warm per-kernel durations, per-stream
busy time, gaps between consecutive kernels of a stream, and how much of the step at least one / two kernels run.
torch.profiler (CUPTI activity records) sees the kernels inside a graph replay with their real start / end times, which
ncu (serialised, cold) cannot give.

  python tools/timeline_probe.py [workload] [lanes on|off] > gpurun_out/timeline.txt
"""
import collections
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

import hlhgat_b200  # noqa: E402
from hlhgat_b200.lib import Hodge_ST_Model as M  # noqa: E402
from hlhgat_b200.parallel import FlatGradBucket  # noqa: E402
from hlhgat_b200.training import Capacity, pad_batch, pad_levels, GraphedTrainStep  # noqa: E402
from hlhgat_b200.workloads import WORKLOADS  # noqa: E402
from hlhgat_b200 import functional as F_hl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "zinc"
lanes = (sys.argv[2] if len(sys.argv) > 2 else "on") == "on"
wl = WORKLOADS[name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
hlhgat_b200.enable_lanes(lanes)
F_hl.enable_factored_hodge1(wl.long_rows)
model = getattr(M, wl.model)(**wl.ctor).to(dev).train()
bucket = FlatGradBucket(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3, fused=True, capturable=True)
raw = [wl.make(wl.batch, i) for i in range(2)]
if wl.levels > 1:
    caps = [Capacity.covering([b[l] for b in raw]) for l in range(wl.levels)]
    host = [pad_levels(b, caps, deg_eps=wl.deg_eps) for b in raw]
else:
    caps = [Capacity.covering(raw)]
    host = [pad_batch(b, caps[0], deg_eps=wl.deg_eps) for b in raw]
stepper = GraphedTrainStep(model, wl.loss, opt, bucket, host[0], dev, warmup=3, loss_fn=True)
for i in range(5):
    stepper.batch.load(host[i % 2])
    stepper.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    stepper.step()
e1.record()
torch.cuda.synchronize()
print(f"workload {name} lanes {'on' if lanes else 'off'}: {e0.elapsed_time(e1) / 10:.3f} ms/step un-profiled (10 replays, same batch)")

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        stepper.step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel" and e.get("ph") == "X"]
ev.sort(key=lambda e: e["ts"])
# the three replays: split at the largest gaps of the fused optimizer kernel (last kernel of a step)
ends = [i for i, e in enumerate(ev) if "multi_tensor_apply" in e["name"] or "FusedOptimizer" in e["name"]]
# take the middle replay: kernels after the last optimizer kernel of step 0 up to the last optimizer kernel of step 1
per = len(ev) // 3
step_ev = ev[per:2 * per]
t0 = min(e["ts"] for e in step_ev)
t1 = max(e["ts"] + e["dur"] for e in step_ev)
print(f"profiled replay: {len(step_ev)} kernels, span {(t1 - t0) / 1e3:.3f} ms")
by_stream = collections.defaultdict(list)
for e in step_ev:
    by_stream[e["args"].get("stream", 0)].append(e)
print("\nper stream: kernels, busy ms, gaps > 0.5 us between consecutive kernels (count, total ms, median us)")
for s, lst in sorted(by_stream.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    busy = sum(e["dur"] for e in lst)
    gaps = [b["ts"] - (a["ts"] + a["dur"]) for a, b in zip(lst, lst[1:])]
    g = sorted(x for x in gaps if x > 0.5)
    med = g[len(g) // 2] if g else 0.0
    print(f"  stream {s}: {len(lst):4d} kernels  busy {busy / 1e3:7.3f} ms  gaps {len(g):4d} / {sum(g) / 1e3:6.3f} ms / median {med:5.2f} us"
          f"  first {(lst[0]['ts'] - t0) / 1e3:6.3f} last {(lst[-1]['ts'] + lst[-1]['dur'] - t0) / 1e3:6.3f}")
# concurrency profile
pts = []
for e in step_ev:
    pts.append((e["ts"], 1))
    pts.append((e["ts"] + e["dur"], -1))
pts.sort()
cur, last, hist = 0, t0, collections.Counter()
for t, d in pts:
    hist[min(cur, 4)] += t - last
    last = t
    cur += d
print("\nconcurrency (kernels running at once): " + ", ".join(f"{k}: {v / 1e3:.3f} ms" for k, v in sorted(hist.items())))
# which kernels run ALONE (nothing overlaps them): candidates for the critical chain / for under-filled SMs
alone, active, last = collections.Counter(), set(), t0
for t, d, i in sorted([(e["ts"], 1, i) for i, e in enumerate(step_ev)] + [(e["ts"] + e["dur"], -1, i) for i, e in enumerate(step_ev)]):
    if t > last and len(active) <= 1:
        alone[step_ev[next(iter(active))]["name"].split("(")[0][:60] if active else "<idle>"] += t - last
    last = t
    if d == 1:
        active.add(i)
    else:
        active.discard(i)
print(f"\ntime with at most one kernel running: {sum(alone.values()) / 1e3:.3f} ms; by kernel:")
for nm, t in alone.most_common(14):
    print(f"{t / 1e3:8.3f} ms  {nm}")
kt, kn = collections.Counter(), collections.Counter()
for e in step_ev:
    nm = e["name"].split("(")[0][:70]
    kt[nm] += e["dur"]
    kn[nm] += 1
tot = sum(kt.values())
print(f"\nkernel time total {tot / 1e3:.3f} ms over {len(step_ev)} kernels (warm, inside the graph, overlapped kernels slow each other)")
for nm, t in kt.most_common(40):
    print(f"{t / 1e3:8.3f} ms {kn[nm]:5d}  {100 * t / tot:5.1f}%  {t / kn[nm]:7.2f} us/launch  {nm}")
if len(sys.argv) > 3:
    out = [(e["name"].split("(")[0][:60], e["args"].get("stream", 0), round(e["ts"] - t0, 2), round(e["dur"], 2)) for e in step_ev]
    with open(sys.argv[3], "w") as f:
        for r in out:
            f.write("\t".join(str(x) for x in r) + "\n")
