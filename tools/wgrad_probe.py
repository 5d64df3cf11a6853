"""Correctness + timing of the tensor-core weight gradient against fp64 and the SIMT split-row kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hlhgat_b200
from hlhgat_b200 import functional as F

dev = "cuda:0"
torch.manual_seed(0)
shapes = [(2048, 64, 64), (24000, 64, 64), (24001, 256, 704), (24000, 128, 320), (24000, 256, 256), (5000, 12, 96)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for R, fo, fi in shapes:
    g, x = torch.randn(R, fo, device=dev), torch.randn(R, fi, device=dev)
    ref = g.double().t() @ x.double()
    out = {}
    tms = {}
    for name in ("tcgen05", "cublas"):
        F.set_dense_backend(name)          # "cublas" here = the SIMT split-row hl_wgrad
        out[name] = F.wgrad(g, x)
        torch.cuda.synchronize()
        for _ in range(3):
            F.wgrad(g, x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            F.wgrad(g, x)
        e1.record()
        torch.cuda.synchronize()
        tms[name] = e0.elapsed_time(e1) / 10 * 1e3
    F.set_dense_backend("tcgen05")
    sc = float(ref.abs().max())
    fl = 2.0 * R * fo * fi
    print(f"R={R} fo={fo} fi={fi}: rel err tcgen05 {float((out['tcgen05'].double() - ref).abs().max()) / sc:.2e} "
          f"simt {float((out['cublas'].double() - ref).abs().max()) / sc:.2e} | tcgen05 {tms['tcgen05']:.1f} us "
          f"({fl / tms['tcgen05'] / 1e6:.1f} TF/s) simt {tms['cublas']:.1f} us ({fl / tms['cublas'] / 1e6:.1f} TF/s)", flush=True)
