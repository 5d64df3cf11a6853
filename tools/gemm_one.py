"""One tcgen05 GEMM shape a few times (target of ncu -k regex:gemm_tf32x3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hlhgat_b200
from hlhgat_b200 import functional as F
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (24000, 256, 704)
a, w, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.randn(N, device="cuda")
g = torch.randn(M, N, device="cuda")
for _ in range(3):
    F.dense(a, w, b)
    F.wgrad(g, a)
torch.cuda.synchronize()
print("ok")
