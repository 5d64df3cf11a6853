"""eigenvector positional encodings (lib/Hodge_Dataset.py:97-112) as a batched GPU
eigensolver.  Eigenvectors are only defined up to sign (and up to a rotation inside a degenerate eigenspace), so parity
is stated on invariants:
  * eigenvalues equal the fp64 LAPACK ones (atol 1e-6 on operators with spectrum in [0, 2]: the iteration runs in fp64,
    only the output is rounded to fp32);
  * every returned column is an eigenvector: ||L v - lambda v|| <= 5e-6, and the columns are orthonormal (1e-5);
  * for an isolated eigenvalue the column equals the reference's up to sign: | |v| - |v_ref| | <= 1e-4;
  * for a cluster of (near-)equal eigenvalues that lies wholly inside the returned range the orthogonal PROJECTOR onto
    its span equals the reference's: || V V^T - V_ref V_ref^T || <= 1e-4.
The reference's own `eig_pe` output (scipy, float32) is in the cache fixtures (tests/golden/cache/expected.pt, columns
21.. of x_t and 3.. of x_s, written by the unmodified reference code)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import hodge_oracle as O


def _clusters(vals, tol):
    groups, start = [], 0
    for i in range(1, len(vals) + 1):
        if i == len(vals) or vals[i] - vals[i - 1] > tol:
            groups.append((start, i))
            start = i
    return groups


def _vec_tol(w, a):
    """An fp32 eigensolver (the reference's scipy call on float32 input) perturbs an eigenvector by ~ eps32 * ||L|| / gap."""
    gap = min(w[a] - w[a - 1] if a > 0 else np.inf, w[a + 1] - w[a] if a + 1 < len(w) else np.inf)
    return 1e-5 + 4e-6 / gap


def check_against_dense(pe, evals, L64, k, gap=2e-3):
    """pe [n, k-1], evals [n] from the GPU against the dense fp64 matrix L64 [n, n]."""
    n = L64.shape[0]
    w, U = np.linalg.eigh(L64)
    assert np.abs(evals - w).max() < 1e-6 * max(1.0, np.abs(w).max()), np.abs(evals - w).max()
    cols = min(k, n) - 1
    V = pe[:, :cols].astype(np.float64)
    assert np.abs(pe[:, cols:]).max(initial=0.0) == 0.0                                 # zero padding for n < k
    if cols <= 0:
        return
    assert np.abs(V.T @ V - np.eye(cols)).max() < 1e-5
    assert np.abs(L64 @ V - V * w[1:cols + 1]).max() < 5e-6 * max(1.0, np.abs(w).max())
    for a, b in _clusters(w, gap):
        lo, hi = max(a, 1), min(b, cols + 1)                                            # ranks of the cluster inside 1 .. cols
        if lo >= hi:
            continue
        if b - a == 1:
            assert np.abs(np.abs(V[:, a - 1]) - np.abs(U[:, a])).max() < 1e-4
        elif a >= 1 and b <= cols + 1:                                                  # whole cluster returned: compare projectors
            P, Pref = V[:, a - 1:b - 1] @ V[:, a - 1:b - 1].T, U[:, a:b] @ U[:, a:b].T
            assert np.abs(P - Pref).max() < 1e-4
    assert (V.max(0) >= -V.min(0) - 1e-6).all()                                         # sign convention: largest |component| positive


def test_oracle_eig_pe_matches_reference_output_in_fixtures():
    """The oracle's `eig_pe` restatement against what the UNMODIFIED reference wrote into the cache fixtures."""
    exp = torch.load(os.path.join(GOLDEN, "cache", "expected.pt"), weights_only=False)
    for g in exp["single"]:
        n, e = g["x_t"].shape[0], g["x_s"].shape[0]
        for x, raw, ei, ew, m in ((g["x_t"], 21, g["edge_index_t"], g["edge_weight_t"], n), (g["x_s"], 3, g["edge_index_s"], g["edge_weight_s"], e)):
            mine = O.eig_pe(O.dense_operator(ei, ew, m), k=100)
            assert mine.shape == x[:, raw:].shape
            L = O.dense_operator(ei, ew, m, torch.float64).numpy()
            w = np.linalg.eigvalsh(L)
            for a, b in _clusters(w, 2e-3):
                if b - a == 1 and a >= 1:                                               # isolated: equal up to sign
                    assert np.abs(np.abs(mine[:, a - 1].numpy()) - np.abs(x[:, raw + a - 1].numpy())).max() < _vec_tol(w, a)


@pytest.mark.gpu
def test_gpu_eig_pe_vs_reference_fixtures_and_lapack():
    from hlhgat_b200.spectral import eig_pe_batch
    from hlhgat_b200.simplex import CsrOperator
    dev = "cuda:0"
    exp = torch.load(os.path.join(GOLDEN, "cache", "expected.pt"), weights_only=False)
    graphs = exp["single"]
    for side, raw in (("t", 21), ("s", 3)):
        off, eis, ews, counts = 0, [], [], []
        for g in graphs:
            m = g["x_" + side].shape[0]
            eis.append(g["edge_index_" + side] + off)
            ews.append(g["edge_weight_" + side])
            counts.append(m)
            off += m
        op = CsrOperator(torch.cat(eis, 1).to(dev), torch.cat(ews).to(dev), off)
        k = 8
        pe, evals = eig_pe_batch(op, counts, k)
        pe, evals = pe.cpu().numpy(), evals.cpu().numpy()
        r = 0
        for g, m in zip(graphs, counts):
            L = O.dense_operator(g["edge_index_" + side], g["edge_weight_" + side], m, torch.float64).numpy()
            check_against_dense(pe[r:r + m], evals[r:r + m], L, k)
            ref = g["x_" + side][:, raw:].numpy()                                       # the reference's own eig_pe columns
            w = np.linalg.eigvalsh(L)
            for a, b in _clusters(w, 2e-3):
                if b - a == 1 and 1 <= a < min(k, m):                               # vs the reference's fp32 LAPACK vectors
                    assert np.abs(np.abs(pe[r:r + m, a - 1]) - np.abs(ref[:, a - 1])).max() < _vec_tol(w, a)
            r += m


@pytest.mark.gpu
@pytest.mark.parametrize("shape,nb,k", [("zinc", 64, 8), ("peptides", 6, 11), ("cifar", 3, 11)])
def test_gpu_eig_pe_on_config_shaped_batches(shape, nb, k):
    """L0 and L1 of config-shaped graphs (ZINC ~23 / 25 rows, peptides ~150, CIFAR superpixels 117 nodes / ~560 edges:
    the L1 blocks there have a ~445-dimensional null space)."""
    from hlhgat_b200.construct import build_simplex_batch
    from hlhgat_b200.spectral import eig_pe_batch
    from hlhgat_b200.synthetic import make_batch
    dev = "cuda:0"
    b = make_batch(shape, nb, seed=3)
    und = b.edge_index
    sb = build_simplex_batch(torch.cat([und[0], und[1]]).to(dev), torch.cat([und[1], und[0]]).to(dev), torch.as_tensor(b.num_node1))
    for op, counts, side in ((sb.op_t, sb.num_node1, "t"), (sb.op_s, sb.num_edge1, "s")):
        pe, evals, vecs, sweeps = eig_pe_batch(op, counts, k, return_all=True)
        assert int(sweeps.max()) < 30, "Jacobi did not converge"
        ei, ew = sb.coo(side)
        ei, ew, pe, evals = ei.cpu(), ew.cpu(), pe.cpu().numpy(), evals.cpu().numpy()
        rows = torch.cat([torch.zeros(1, dtype=torch.long), torch.as_tensor(counts).cpu().cumsum(0)])
        for g in range(min(nb, 4)):
            r0, r1 = int(rows[g]), int(rows[g + 1])
            sel = (ei[0] >= r0) & (ei[0] < r1)
            L = O.dense_operator(ei[:, sel] - r0, ew[sel], r1 - r0, torch.float64).numpy()
            check_against_dense(pe[r0:r1], evals[r0:r1], L, k)
            Vall = vecs[g].cpu().double().numpy()
            assert np.abs(Vall.T @ Vall - np.eye(r1 - r0)).max() < 2e-5
            assert np.abs(Vall @ np.diag(evals[r0:r1]) @ Vall.T - L).max() < 1e-5           # L = V diag(lambda) V^T


@pytest.mark.gpu
def test_gpu_eig_pe_group_fc_spectrum():
    """The DEMO's brain skeleton (N = 268, max degree 142): low and high end of the L0 spectrum vs the fixture, and the
    reference-signature single-matrix call."""
    from hlhgat_b200.spectral import eig_pe
    g = load_golden("group_fc.pt")
    n = g["num_nodes"]
    L = O.dense_operator(g["ei_t"].long(), g["w_t"], n)
    pe = eig_pe(L.to("cuda:0"), k=10)
    assert pe.shape == (n, 9)
    from hlhgat_b200.spectral import eig_pe_batch
    from hlhgat_b200.simplex import CsrOperator
    op = CsrOperator(g["ei_t"].long().to("cuda:0"), g["w_t"].to("cuda:0"), n)
    pe2, evals = eig_pe_batch(op, [n], 10)
    assert torch.equal(pe, pe2)
    ev = evals.cpu().double()
    assert float((ev[:24] - g["l0_spectrum_low"]).abs().max()) < 2e-6 and float((ev[-8:] - g["l0_spectrum_high"]).abs().max()) < 2e-6
    check_against_dense(pe.cpu().numpy(), evals.cpu().numpy(), L.double().numpy(), 10)
