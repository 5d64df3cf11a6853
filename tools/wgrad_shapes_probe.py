"""Per-shape device time of the tcgen05 3xTF32 weight gradient dW[fo,fi] = g[R,fo]^T x[R,fi] (+ its split reduce) on
every weight-gradient shape of the ZINC bench model, inside a CUDA graph (20 launches per replay)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hlhgat_b200  # noqa: E402,F401
from hlhgat_b200 import _native as N  # noqa: E402

L = N.lib()
dev = "cuda:0"
R = 24144


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


def run(fo, fi, label, count):
    g, x = torch.randn(R, fo, device=dev), torch.randn(R, fi, device=dev)
    dw = torch.empty(fo, fi, device=dev)
    nb = L.hl_wgrad_tf32x3_workspace(R, fo, fi)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)

    def fn():
        rc = L.hl_wgrad_tf32x3(g.data_ptr(), fo, x.data_ptr(), fi, R, fo, fi, dw.data_ptr(), fi, 0, ws.data_ptr(), nb, N.stream_ptr())
        assert rc == 0, rc
    us = timed(fn)
    ref = g.double().t() @ x.double()
    err = float((dw.double() - ref).abs().max() / ref.abs().max())
    fl = 2.0 * R * fo * fi
    print(f"{label:18s} fo={fo:4d} fi={fi:4d} (x{count}): {us:7.1f} us  {fl / us / 1e6:6.1f} TFLOP/s  rel err {err:.1e}  ws {nb / 1e6:.1f} MB", flush=True)
    return us * count


tot = 0.0
for d, f in ((64, 64), (128, 64), (192, 128), (320, 128), (448, 256), (704, 256)):
    tot += run(f, d, "wgrad MLP1 halves", 2)
    tot += run(f, f, "wgrad MLP2 + conv", 3)
print(f"sum over one side of the model: {tot / 1e3:.2f} ms (x2 sides per step)")
