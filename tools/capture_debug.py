"""Debug aid: capture the small ZINC training step with / without the dense stack and print the first CUDA error with
C++ frames (run with TORCH_SHOW_CPP_STACKTRACES=1)."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hlhgat_b200  # noqa: E402
from hlhgat_b200.dense_stack import enable_dense_stack  # noqa: E402
from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr  # noqa: E402
from hlhgat_b200.parallel import FlatGradBucket  # noqa: E402
from hlhgat_b200.synthetic import make_batch  # noqa: E402
from hlhgat_b200.training import Capacity, pad_batch, GraphedTrainStep  # noqa: E402

DEV = "cuda:0"
CTOR = dict(channels=[1, 2], filters=[32, 64], mlp_channels=[48], K=3, node_dim=21, edge_dim=3, keig=7)
for stack in (False, True):
    enable_dense_stack(stack)
    torch.manual_seed(0)
    raws = [make_batch("zinc", 32, seed=s) for s in (3, 4)]
    cap = Capacity.covering(raws, slack=0.1)
    host = [pad_batch(r, cap, pin=True) for r in raws]
    m = HL_HGCNN_zinc_dense_int3_pyr(**CTOR).to(DEV).train()
    try:
        st = GraphedTrainStep(m, torch.nn.L1Loss(), torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True),
                              FlatGradBucket(m.parameters()), host[0], DEV, warmup=int(os.environ.get("WARMUP", "3")))
        print("stack", stack, "captured ok, loss", float(st.step()))
    except Exception:
        print("stack", stack, "FAILED")
        traceback.print_exc()
        break
