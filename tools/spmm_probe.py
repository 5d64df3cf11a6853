"""Launch the polynomial SpMM (fused L0+L1) a few times on a > L2 ZINC-shaped operator stack, or on a
TSP-shaped L1 -- the target of `ncu --set full -k regex:poly_spmm`.  Prints CUDA-event timings."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hlhgat_b200  # noqa: E402,F401
from hlhgat_b200 import functional as F_hl, _native as N  # noqa: E402
from hlhgat_b200.simplex import CsrOperator  # noqa: E402
from hlhgat_b200.synthetic import make_batch, batch_to  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="zinc")
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--reps", type=int, default=16)
ap.add_argument("--width", type=int, default=64)
ap.add_argument("--K", type=int, default=2)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--family", default="laguerre")
ap.add_argument("--factored", action="store_true", help="edge operator only, applied as diag(2/lambda) B1^T (B1 x): hodge1_* kernels")
a = ap.parse_args()
dev = torch.device("cuda:0")
if a.shape == "tsp":
    from hlhgat_b200.synthetic import make_tsp_batch  # noqa: E402
    b = batch_to(make_tsp_batch(a.batch, seed=0), dev)
else:
    b = batch_to(make_batch(a.shape, a.batch, seed=0), dev)
ops, xs, nnz, rows = [], [], 0, 0
for ei, ew, r in ((b.edge_index_t, b.edge_weight_t, b.x_t.shape[0]), (b.edge_index_s, b.edge_weight_s, b.x_s.shape[0])):
    off = (torch.arange(a.reps, device=dev) * r).repeat_interleave(ei.shape[1])
    op = CsrOperator(ei.repeat(1, a.reps) + off, ew.repeat(a.reps), r * a.reps)
    op.fwd
    ops.append(op)
    xs.append(torch.randn(r * a.reps, a.width, device=dev))
    nnz += ei.shape[1] * a.reps
    rows += r * a.reps
if a.factored:
    from hlhgat_b200.simplex import Hodge1Factor, incidence_for  # noqa: E402
    n0, e0 = b.x_t.shape[0], b.x_s.shape[0]
    off = (torch.arange(a.reps, device=dev) * n0).repeat_interleave(e0)
    inc = incidence_for(b.edge_index.repeat(1, a.reps) + off, n0 * a.reps)
    ops[1].factored = Hodge1Factor.from_operator(ops[1], inc)
    F_hl.enable_factored_hodge1(True)
    ops, xs = [ops[1]], [xs[1]]
    n, e = n0 * a.reps, e0 * a.reps
    rows = e
    nnz = 0
fam = N.HL_LAGUERRE if a.family == "laguerre" else N.HL_CHEB
for _ in range(2):
    F_hl.poly_basis_fwd(fam, a.K, ops, xs, a.width)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    F_hl.poly_basis_fwd(fam, a.K, ops, xs, a.width)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters / max(a.K - 1, 1)
first = 8 * nnz + 4 * (rows + 2) + 4 * rows * a.width * 2
if a.factored:          # node pass + edge pass of the factored form, n_epi = 1
    first = 4 * (n + 1) + 16 * e + 4 * a.width * (e + n) + 12 * e + 4 * a.width * (n + 2 * e)
print(f"shape={a.shape} rows={rows} nnz={nnz} F={a.width} K={a.K}: {ms * 1e3:.1f} us/launch, "
      f"first-order algorithmic {first / 1e6:.1f} MB -> {first / ms / 1e6:.0f} GB/s")
