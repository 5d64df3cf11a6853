import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hlhgat_b200
from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
from hlhgat_b200.synthetic import make_batch, batch_to
from oracle import hodge_oracle as O
ctor = dict(channels=[1, 1], filters=[32, 64], mlp_channels=[], K=3, node_dim=21, edge_dim=3, keig=7)
for ng in (16, 64):
    torch.manual_seed(0)
    ref = O.HL_HGCNN_zinc_dense_int3_pyr(**ctor).train()
    b = make_batch("zinc", ng, seed=0)
    pr = ref(b); torch.nn.functional.l1_loss(pr, b.y).backward()
    for lanes in (False, True):
        model = HL_HGCNN_zinc_dense_int3_pyr(**ctor).to("cuda:0").train()
        model.load_state_dict(ref.state_dict(), strict=True)
        d = batch_to(b, "cuda:0")
        hlhgat_b200.enable_lanes(lanes)
        pred = model(d, device="cuda:0"); torch.nn.functional.l1_loss(pred, d.y).backward()
        hlhgat_b200.enable_lanes(False)
        torch.cuda.synchronize()
        worst = max(((p.grad.cpu() - q.grad).norm() / (q.grad.norm() + 1e-12), n) for (n, p), q in zip(model.named_parameters(), ref.parameters()) if n.endswith("weight"))
        ga = model.NEConv00.module_4.lins[2].weight.grad.cpu(); gr = ref.NEConv00.module_4.lins[2].weight.grad
        print(f"graphs {ng} lanes {lanes}: pred err {float((pred.cpu()-pr).abs().max()):.2e}  lins[2] rel {float((ga-gr).norm()/gr.norm()):.2e}  worst weight {float(worst[0]):.2e} {worst[1]}", flush=True)
