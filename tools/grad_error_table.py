"""Per-parameter gradient error of the CUDA path and of the fp32 CPU oracle, both against the fp64 oracle."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hlhgat_b200
from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
from hlhgat_b200.synthetic import make_batch, batch_to
from oracle import hodge_oracle as O
torch.manual_seed(0)
ctor = dict(channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[], K=2, node_dim=21, edge_dim=3, keig=7)
ref = O.HL_HGCNN_zinc_dense_int3_pyr(**ctor).train()
b = make_batch("zinc", int(sys.argv[1]) if len(sys.argv) > 1 else 64, seed=11)
g_ref = torch.autograd.grad(torch.nn.functional.l1_loss(ref(b), b.y), list(ref.parameters()), allow_unused=True)
ref64 = copy.deepcopy(ref).double()
b64 = copy.copy(b)
for k in ("x_t", "x_s", "y", "edge_weight_t", "edge_weight_s"):
    setattr(b64, k, getattr(b, k).double())
g64 = torch.autograd.grad(torch.nn.functional.l1_loss(ref64(b64), b64.y), list(ref64.parameters()), allow_unused=True)
model = HL_HGCNN_zinc_dense_int3_pyr(**ctor).to("cuda:0").train()
model.load_state_dict(ref.state_dict())
d = batch_to(b, "cuda:0")
g = torch.autograd.grad(torch.nn.functional.l1_loss(model(d, device="cuda:0"), d.y), list(model.parameters()), allow_unused=True)
for (n, _), a, r, r64 in zip(model.named_parameters(), g, g_ref, g64):
    if r is None:
        continue
    s = float(r64.norm())
    print(f"{n:45s} scale {s:9.3e} ours {float((a.cpu().double() - r64).norm()) / (s + 1e-30):9.2e} cpu32 {float((r.double() - r64).norm()) / (s + 1e-30):9.2e}")
