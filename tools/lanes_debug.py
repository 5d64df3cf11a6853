"""Localise any difference between single-stream and two-lane issue: forward outputs of every block, then every
parameter gradient, compared bit for bit (off/off, on/on, off/on)."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hlhgat_b200 as H  # noqa: E402
from hlhgat_b200.lib import Hodge_ST_Model as M  # noqa: E402
from hlhgat_b200.synthetic import make_batch, batch_to  # noqa: E402

DEV = "cuda:0"
CTOR = dict(channels=[1, 2], filters=[32, 64], mlp_channels=[48], K=3, node_dim=21, edge_dim=3, keig=7)
torch.manual_seed(0)
b = batch_to(make_batch("zinc", 96, seed=7), DEV)
base = M.HL_HGCNN_zinc_dense_int3_pyr(**CTOR).to(DEV).train()


def run(use):
    H.enable_lanes(use)
    m = copy.deepcopy(base)
    acts = {}

    def hook(name):
        def fn(mod, inp, out):
            outs = out if isinstance(out, (list, tuple)) else [out]
            for i, o in enumerate(outs):
                if torch.is_tensor(o):
                    acts[f"{name}[{i}]"] = o
        return fn
    for name, mod in m.named_modules():
        if name and name.count(".") <= 1:
            mod.register_forward_hook(hook(name))
    loss = torch.nn.functional.l1_loss(m(b, device=DEV), b.y)
    names = list(acts)
    grads_act = torch.autograd.grad(loss, [acts[n] for n in names] + list(m.parameters()), allow_unused=True)
    torch.cuda.synchronize()
    H.enable_lanes(False)
    res = {"loss": loss.detach().clone()}
    for n in names:
        res["act " + n] = acts[n].detach().clone()
    for n, g in zip(names + [k for k, _ in m.named_parameters()], grads_act):
        if g is not None:
            res["grad " + n] = g.clone()
    return res


def diff(a, c, label):
    bad = [(k, float((a[k] - c[k]).abs().max()), float(a[k].abs().max())) for k in a if k in c and not torch.equal(a[k], c[k])]
    print(f"== {label}: {len(bad)} of {len(a)} tensors differ", flush=True)
    for k, d, s in bad[:60]:
        print(f"   {k:60s} max|diff| {d:.3e}  (scale {s:.3e})")


off1, off2 = run(False), run(False)
on1, on2 = run(True), run(True)
diff(off1, off2, "off vs off")
diff(on1, on2, "on vs on")
diff(off1, on1, "off vs on")
