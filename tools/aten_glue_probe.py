"""Which ATen ops (and from where) launch kernels inside one eager training step of the bench model:
torch.profiler with python stacks, grouped by op and innermost hlhgat_b200 frame."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

import hlhgat_b200  # noqa: E402,F401
from hlhgat_b200.lib import Hodge_ST_Model as M  # noqa: E402
from hlhgat_b200.parallel import FlatGradBucket  # noqa: E402
from hlhgat_b200.training import Capacity, pad_batch, StaticBatch  # noqa: E402
from hlhgat_b200.workloads import WORKLOADS  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "zinc"]
dev = torch.device("cuda:0")
model = getattr(M, wl.model)(**wl.ctor).to(dev).train()
bucket = FlatGradBucket(model.parameters())
from hlhgat_b200.training import pad_levels  # noqa: E402
raw = [wl.make(wl.batch, 0)]
if wl.levels > 1:
    caps = [Capacity.covering([raw[0][l]]) for l in range(wl.levels)]
    sb = StaticBatch(pad_levels(raw[0], caps, deg_eps=wl.deg_eps), dev)
else:
    cap = Capacity.covering(raw)
    sb = StaticBatch(pad_batch(raw[0], cap, deg_eps=wl.deg_eps), dev)


from hlhgat_b200.functional import accumulate_into_grads  # noqa: E402


def step():
    bucket.zero()
    loss = wl.loss(model, sb)
    with accumulate_into_grads():
        loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
by = collections.Counter()
for ev in prof.events():
    if not ev.name.startswith("aten::") or ev.device_time_total <= 0 or not ev.kernels:
        continue
    where = "?"
    for fr in ev.stack:
        if "hl-hgat_b200" in fr or "workloads" in fr:
            where = fr.split("hl-hgat_b200/")[-1]
            break
    if where == "?" and ev.stack:
        where = "autograd:" + "|".join(s.split("/")[-1] for s in ev.stack[:2])
    by[(ev.name, where)] += len(ev.kernels)
for (name, where), n in sorted(by.items(), key=lambda kv: -kv[1])[:60]:
    print(f"{n:5d}  {name:28s} {where}")

# real (warm, un-profiled-by-ncu) kernel durations of the same step
kt = collections.Counter()
kn = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        kt[ev.name[:60]] += ev.device_time_total
        kn[ev.name[:60]] += 1
tot = sum(kt.values())
print(f"kernel time total {tot / 1e3:.2f} ms over {sum(kn.values())} kernels")
for name, t in kt.most_common(45):
    print(f"{t / 1e3:8.3f} ms {kn[name]:5d}  {100 * t / tot:5.1f}%  {name}")
