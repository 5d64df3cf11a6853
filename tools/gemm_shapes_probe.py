"""Per-shape device time of the tcgen05 3xTF32 GEMM (forward y = x W^T and data gradient g W) on every dense
shape of the ZINC bench model, measured inside a CUDA graph (20 launches per replay: no host launch gaps).
Prints us per launch, TFLOP/s and the HBM floor (fp32 operands + output read/written once at the measured peak)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hlhgat_b200  # noqa: E402,F401
from hlhgat_b200 import _native as N  # noqa: E402

L = N.lib()
dev = "cuda:0"
R = 24144
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


def run(M, Nn, K, label):
    a = torch.randn(M, K, device=dev)
    w = torch.randn(Nn, K, device=dev) * 0.1
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    N.check(L.hl_tf32_split(w.data_ptr(), w.stride(0), Nn, K, 0, hi.data_ptr(), lo.data_ptr(), K, N.stream_ptr()), "split")
    c = torch.empty(M, Nn, device=dev)

    def fn():
        rc = L.hl_gemm_tf32x3(a.data_ptr(), a.stride(0), hi.data_ptr(), lo.data_ptr(), K, M, Nn, K, None, c.data_ptr(), c.stride(0), 0,
                              N.stream_ptr())
        assert rc == 0, rc
    us = timed(fn)
    fl = 2.0 * M * Nn * K
    floor = 4.0 * (M * K + M * Nn + 2 * Nn * K) / HBM / 1e3
    print(f"{label:22s} M={M} N={Nn:4d} K={K:5d}: {us:7.1f} us  {fl / us / 1e6:6.1f} TFLOP/s   HBM floor {floor:5.1f} us", flush=True)
    return us


tot = 0.0
for d, f in ((64, 64), (128, 64), (192, 128), (320, 128), (448, 256), (704, 256)):
    tot += run(R, f, 2 * d, "fwd MLP1")
    tot += run(R, f, f, "fwd MLP2")
    tot += run(R, f, 2 * f, "fwd conv K=2")
    tot += 2 * run(R, d, f, "dgrad MLP1 (x2)")
    tot += run(R, f, f, "dgrad MLP2")
    tot += 2 * run(R, f, f, "dgrad conv (x2)")
print(f"sum over one side of the model: {tot / 1e3:.2f} ms (x2 sides per step)")
