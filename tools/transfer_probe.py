"""Node <-> edge transfers on a > L2 ZINC-shaped incidence stack (bench batch x reps, block-diagonal): s2t = (1/D)|B1| x_s
(segment_reduce_kernel), t2s = |B1|^T x_t / 2 (endpoint_gather_kernel) and the adjoint of t2s (segment_reduce, constant
post-scale).  Prints CUDA-event timings and GB/s of ALGORITHMIC bytes; the target of
`ncu --set full -k regex:segment_reduce|endpoint_gather`."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hlhgat_b200  # noqa: E402,F401
from hlhgat_b200 import functional as F_hl, _native as N  # noqa: E402
from hlhgat_b200.simplex import incidence_for  # noqa: E402
from hlhgat_b200.synthetic import make_batch, batch_to  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="zinc")
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--reps", type=int, default=16)
ap.add_argument("--widths", default="64,256")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
b = batch_to(make_batch(a.shape, a.batch, seed=0), dev)
n0, e0 = b.x_t.shape[0], b.x_s.shape[0]
off = (torch.arange(a.reps, device=dev) * n0).repeat_interleave(e0)
inc = incidence_for(b.edge_index.repeat(1, a.reps) + off, n0 * a.reps)
n, e = n0 * a.reps, e0 * a.reps
D = inc.degree() + 1e-6
for width in [int(w) for w in a.widths.split(",")]:
    x_s, x_t = torch.randn(e, width, device=dev), torch.randn(n, width, device=dev)
    cases = {
        "s2t  (1/D)|B1| x_s      segment_reduce": (lambda: F_hl._segment_reduce(inc.rowptr, inc.edge, n, x_s, N.HL_POST_RCP_ROW, row_scale=D),
                                                   4 * (n + 1) + 4 * 2 * e + 4 * n + 4 * width * (e + n)),
        "t2s  |B1|^T x_t / 2     endpoint_gather": (lambda: F_hl._endpoint_gather(inc, x_t, None, 0.5), 8 * e + 4 * width * (n + e)),
        "t2s^T (adjoint)         segment_reduce": (lambda: F_hl._segment_reduce(inc.rowptr, inc.edge, n, x_s, N.HL_POST_CONST, cscale=0.5),
                                                   4 * (n + 1) + 4 * 2 * e + 4 * width * (e + n)),
    }
    for name, (fn, alg) in cases.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0_.record()
        for _ in range(a.iters):
            fn()
        e1_.record()
        torch.cuda.synchronize()
        ms = e0_.elapsed_time(e1_) / a.iters
        print(f"F={width:4d} {name}: N={n} E={e}: {ms * 1e3:7.1f} us/launch, algorithmic {alg / 1e6:7.1f} MB -> {alg / ms / 1e6:6.0f} GB/s")
