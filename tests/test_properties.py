"""Property tests: random simple graphs -> nnz identities, symmetry of L0 / L1, zero row sums of L0,
lambda_max(L0) = lambda_max(L1), spectrum of the normalised operators inside [0, 2], block-diagonal batching = per-graph
results concatenated.  On the CPU for the oracle's construction; on the GPU for hl_build_edges / hl_lambda_max /
hl_laplacian_* and the polynomial convolution."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st, HealthCheck

from oracle import hodge_oracle as O

SET = dict(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))


@st.composite
def simple_graph(draw, n_max=14):
    n = draw(st.integers(2, n_max))
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    k = draw(st.integers(1, min(len(pairs), 3 * n)))
    idx = draw(st.lists(st.integers(0, len(pairs) - 1), min_size=k, max_size=k, unique=True))
    und = np.array([pairs[i] for i in sorted(idx)]).T
    flip = draw(st.lists(st.booleans(), min_size=und.shape[1], max_size=und.shape[1]))
    a = np.where(flip, und[1], und[0])
    b = np.where(flip, und[0], und[1])
    ei = np.concatenate([np.stack([a, b]), np.stack([b, a])], 1)            # both directions, arbitrary orientation first
    perm = draw(st.permutations(range(ei.shape[1])))
    return n, torch.from_numpy(ei[:, list(perm)].copy()).long()


def _dense(ei, w, n):
    return torch.zeros(n, n, dtype=torch.float64).index_put_((ei[0], ei[1]), w.double(), accumulate=True)


def _check_operator_properties(n, e, und, ei_t, w_t, ei_s, w_s, lam):
    deg = torch.zeros(n).index_add_(0, und.reshape(-1), torch.ones(2 * e))
    assert ei_t.shape[1] == int((deg > 0).sum()) + 2 * e                       # no diagonal entry for isolated nodes
    assert ei_s.shape[1] == e + int((deg * (deg - 1)).sum())
    L0, L1 = _dense(ei_t, w_t, n), _dense(ei_s, w_s, e)
    assert torch.equal(L0, L0.T) and torch.equal(L1, L1.T)
    assert float(L0.sum(1).abs().max()) < 1e-6
    ev0, ev1 = torch.linalg.eigvalsh(L0), torch.linalg.eigvalsh(L1)
    assert abs(float(ev0.max()) - 2.0) < 1e-5 and abs(float(ev1.max()) - 2.0) < 1e-5     # both scaled by 2 / lambda_max
    assert float(ev0.min()) > -1e-6 and float(ev1.min()) > -1e-6
    B1 = O.adj2par1(und, n, e, torch.float64).to_dense()
    assert abs(float(torch.linalg.eigvalsh(B1 @ B1.T).max()) - float(lam)) < 1e-5 * max(1.0, float(lam))
    # sign rule of the off-diagonals of B1^T B1 and L0 = D - A
    assert torch.allclose(L0 * float(lam) / 2, B1 @ B1.T, atol=1e-5) and torch.allclose(L1 * float(lam) / 2, B1.T @ B1, atol=1e-5)


@settings(**SET)
@given(simple_graph())
def test_oracle_construction_properties(g):
    n, ei = g
    s = O.build_simplex_graph(ei, n)
    und = s.edge_index
    assert bool((und[0] < und[1]).all()) and torch.equal(und, und[:, torch.argsort(und[0] * n + und[1])])
    _check_operator_properties(n, und.shape[1], und, s.edge_index_t, s.edge_weight_t, s.edge_index_s, s.edge_weight_s, s.maxeig)


@settings(**SET)
@given(st.lists(simple_graph(10), min_size=2, max_size=4), st.integers(1, 4), st.sampled_from(["laguerre", "cheb"]))
def test_oracle_block_diagonal_batching_equals_per_graph(graphs, K, family):
    torch.manual_seed(0)
    cls = O.HodgeLaguerreConv if family == "laguerre" else O.HodgeChebConv
    conv = cls(3, 5, K)
    parts, outs = [], []
    for n, ei in graphs:
        s = O.build_simplex_graph(ei, n)
        s.x_t, s.x_s = torch.randn(n, 3), torch.randn(s.edge_index.shape[1], 3)
        parts.append(s)
        outs.append((conv(s.x_t, s.edge_index_t, s.edge_weight_t), conv(s.x_s, s.edge_index_s, s.edge_weight_s)))
    b = O.collate(parts)
    yt = conv(b.x_t, b.edge_index_t, b.edge_weight_t)
    ys = conv(b.x_s, b.edge_index_s, b.edge_weight_s)
    assert torch.allclose(yt, torch.cat([o[0] for o in outs]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(ys, torch.cat([o[1] for o in outs]), rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@settings(**SET)
@given(st.lists(simple_graph(12), min_size=1, max_size=5))
def test_gpu_construction_properties_and_batching(graphs):
    from hlhgat_b200.construct import build_simplex_batch
    dev = "cuda:0"
    off, src, dst = 0, [], []
    for n, ei in graphs:
        src.append(ei[0] + off)
        dst.append(ei[1] + off)
        off += n
    sb = build_simplex_batch(torch.cat(src).to(dev), torch.cat(dst).to(dev), torch.tensor([n for n, _ in graphs]))
    ei_t, w_t = (t.cpu() for t in sb.coo("t"))
    ei_s, w_s = (t.cpu() for t in sb.coo("s"))
    n_off = e_off = t_off = s_off = 0
    for k, (n, ei) in enumerate(graphs):
        s = O.build_simplex_graph(ei, n)                                            # per-graph oracle
        e, nt, ns = s.edge_index.shape[1], s.edge_index_t.shape[1], s.edge_index_s.shape[1]
        assert torch.equal(sb.edge_index[:, e_off:e_off + e].cpu() - n_off, s.edge_index)
        assert torch.equal(ei_t[:, t_off:t_off + nt] - n_off, s.edge_index_t)      # block-diagonal batch = graphs concatenated
        assert torch.equal(ei_s[:, s_off:s_off + ns] - e_off, s.edge_index_s)
        assert torch.allclose(w_t[t_off:t_off + nt], s.edge_weight_t, rtol=2e-6, atol=0)
        assert torch.allclose(w_s[s_off:s_off + ns], s.edge_weight_s, rtol=2e-6, atol=0)
        _check_operator_properties(n, e, s.edge_index, ei_t[:, t_off:t_off + nt] - n_off, w_t[t_off:t_off + nt],
                                   ei_s[:, s_off:s_off + ns] - e_off, w_s[s_off:s_off + ns], sb.lambda_max[k].cpu())
        n_off, e_off, t_off, s_off = n_off + n, e_off + e, t_off + nt, s_off + ns
    assert ei_t.shape[1] == t_off and ei_s.shape[1] == s_off


@pytest.mark.gpu
@settings(**SET)
@given(st.lists(simple_graph(10), min_size=2, max_size=4), st.integers(1, 4), st.sampled_from(["laguerre", "cheb"]),
       st.sampled_from([3, 8, 20]))
def test_gpu_conv_batching_equals_per_graph_bit_exact(graphs, K, family, width):
    """Rows of a block-diagonal operator only see their own graph: the batched polynomial basis equals the per-graph
    bases concatenated BIT FOR BIT (same per-row summation order)."""
    from hlhgat_b200 import functional as F_hl, _native as N
    from hlhgat_b200.simplex import CsrOperator
    dev = "cuda:0"
    fam = N.HL_LAGUERRE if family == "laguerre" else N.HL_CHEB
    torch.manual_seed(1)
    parts, bases = [], []
    for n, ei in graphs:
        s = O.build_simplex_graph(ei, n)
        s.x_t, s.x_s = torch.randn(n, width), torch.randn(s.edge_index.shape[1], width)
        op = CsrOperator(s.edge_index_s.to(dev), s.edge_weight_s.to(dev), s.x_s.shape[0])
        bases.append(F_hl.poly_basis_fwd(fam, K + 1, [op], [s.x_s.to(dev)], width)[0])
        parts.append(s)
    b = O.collate(parts)
    op = CsrOperator(b.edge_index_s.to(dev), b.edge_weight_s.to(dev), b.x_s.shape[0])
    whole = F_hl.poly_basis_fwd(fam, K + 1, [op], [b.x_s.to(dev)], width)[0]
    assert torch.equal(whole, torch.cat(bases, dim=1))
