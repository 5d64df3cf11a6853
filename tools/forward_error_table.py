"""Per-module forward error of the CUDA model vs the fp64 oracle (same weights, same batch)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hlhgat_b200
from hlhgat_b200 import functional as F
from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
from hlhgat_b200.synthetic import make_batch, batch_to
from oracle import hodge_oracle as O
if len(sys.argv) > 2:
    F.set_dense_backend(sys.argv[2])
torch.manual_seed(0)
ctor = dict(channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[], K=2, node_dim=21, edge_dim=3, keig=7)
ref = O.HL_HGCNN_zinc_dense_int3_pyr(**ctor).train()
b = make_batch("zinc", int(sys.argv[1]) if len(sys.argv) > 1 else 64, seed=11)
ref64 = copy.deepcopy(ref).double()
ref32 = copy.deepcopy(ref)
b64 = copy.copy(b)
for k in ("x_t", "x_s", "y", "edge_weight_t", "edge_weight_s"):
    setattr(b64, k, getattr(b, k).double())
model = HL_HGCNN_zinc_dense_int3_pyr(**ctor).to("cuda:0").train()
model.load_state_dict(ref.state_dict())
outs = {"ours": {}, "f64": {}, "f32": {}}


def hook(store, name):
    def fn(mod, inp, out):
        o = out if torch.is_tensor(out) else out[0]
        store[name] = o.detach().double().cpu()
    return fn


for tag, m in (("ours", model), ("f64", ref64), ("f32", ref32)):
    for name, mod in m.named_modules():
        if name and name.count(".") <= 1:
            mod.register_forward_hook(hook(outs[tag], name))
model(batch_to(b, "cuda:0"), device="cuda:0")
ref64(b64)
ref32(b)
for name in outs["f64"]:
    if name in outs["ours"] and outs["ours"][name].shape == outs["f64"][name].shape:
        r = outs["f64"][name]
        s = float(r.abs().max())
        eo = float((outs["ours"][name] - r).abs().max()) / s
        ec = float((outs["f32"][name] - r).abs().max()) / s
        print(f"{name:32s} max|ref| {s:9.3e}  ours {eo:9.2e}  cpu32 {ec:9.2e}")
