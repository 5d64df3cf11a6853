import sys, os
sys.path.insert(0, "/root/repo")
import torch, copy
import hlhgat_b200
from hlhgat_b200.lib import Hodge_ST_Model as M
from hlhgat_b200.parallel import FlatGradBucket, FlatAdam
from hlhgat_b200.training import Capacity, pad_batch, GraphedTrainStep
from hlhgat_b200.workloads import WORKLOADS
wl = WORKLOADS["zinc"]; dev = "cuda:0"
raws = [wl.make(256, i) for i in range(3)]
cap = Capacity.covering(raws); host = [pad_batch(r, cap, pin=True) for r in raws]
hlhgat_b200.enable_lanes(True)
out = {}
for kind in ("torch", "flat"):
    torch.manual_seed(0)
    model = getattr(M, wl.model)(**wl.ctor).to(dev).train()
    bucket = FlatGradBucket(model.parameters())
    opt = FlatAdam(bucket, lr=1e-3, weight_decay=1e-3) if kind == "flat" else torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3, fused=True, capturable=True)
    st = GraphedTrainStep(model, wl.loss, opt, bucket, host[0], dev, warmup=3, loss_fn=True)
    ls = []
    for i in range(12):
        st.batch.load(host[i % 3]); ls.append(float(st.step()))
    out[kind] = ls
    print(kind, " ".join(f"{v:.5f}" for v in ls))
print("max rel diff", max(abs(a - b) / abs(a) for a, b in zip(out["torch"], out["flat"])))
