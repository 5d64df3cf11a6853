"""Correctness + timing of the tcgen05 3xTF32 GEMM against fp64 and cuBLAS fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hlhgat_b200
from hlhgat_b200 import _native as N

L = N.lib()
dev = "cuda:0"


def gemm(a, w, bias=None, out=None, accumulate=False):
    M, K = a.shape
    Nn = w.shape[0]
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    N.check(L.hl_tf32_split(w.data_ptr(), w.stride(0), Nn, K, 0, hi.data_ptr(), lo.data_ptr(), K, N.stream_ptr()), "split")
    c = out if out is not None else torch.empty(M, Nn, device=dev)
    rc = L.hl_gemm_tf32x3(a.data_ptr(), a.stride(0), hi.data_ptr(), lo.data_ptr(), K, M, Nn, K, N.ptr(bias), c.data_ptr(),
                          c.stride(0), 1 if accumulate else 0, N.stream_ptr())
    assert rc == 0, rc
    return c


torch.manual_seed(0)
shapes = [(256, 64, 64), (1000, 128, 96), (24000, 64, 64), (24000, 256, 1408), (24000, 128, 28), (24001, 256, 704), (3000, 32, 64)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for M, Nn, K in shapes:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(Nn, K, device=dev) * 0.1
    bias = torch.randn(Nn, device=dev)
    c = gemm(a, w, bias)
    torch.cuda.synchronize()
    ref64 = (a.double() @ w.double().t() + bias.double())
    ref32 = torch.addmm(bias, a, w.t())
    e_ours = float((c.double() - ref64).abs().max() / ref64.abs().max())
    e_cublas = float((ref32.double() - ref64).abs().max() / ref64.abs().max())
    c2 = gemm(a, w, None, out=c.clone(), accumulate=True)
    e_acc = float((c2.double() - (2 * ref64 - bias.double())).abs().max() / ref64.abs().max())
    for _ in range(3):
        gemm(a, w, bias); torch.addmm(bias, a, w.t())
    torch.cuda.synchronize()
    t = []
    for fn in (lambda: gemm(a, w, bias), lambda: torch.addmm(bias, a, w.t())):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1) / 10 * 1e3)
    fl = 2.0 * M * Nn * K
    print(f"M={M} N={Nn} K={K}: rel err ours {e_ours:.2e} (acc {e_acc:.2e}) cublas-fp32 {e_cublas:.2e} | "
          f"ours {t[0]:.1f} us ({fl / t[0] / 1e6:.1f} TFLOP/s)  cublas {t[1]:.1f} us ({fl / t[1] / 1e6:.1f} TFLOP/s)", flush=True)
