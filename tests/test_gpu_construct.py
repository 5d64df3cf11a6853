"""GPU simplex-graph construction against the reference's construction (golden vectors recorded
from lib/Hodge_Dataset.py formulas) and the CPU oracle: indices and signs bit-exact, values
within 2 ulp (lambda_max comes from fp64 Lanczos instead of fp32 LAPACK eigh)."""
import numpy as np
import pytest
import torch

import hlhgat_b200  # noqa: F401
from hlhgat_b200.construct import build_simplex_batch
from hlhgat_b200.synthetic import make_batch, SHAPES, _knn_graph, _tree_plus_chords
from oracle import hodge_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _check_against(graphs, sb, rtol=3e-7):
    n_off = e_off = t_off = s_off = 0
    ei_t, ew_t = sb.coo("t")
    ei_s, ew_s = sb.coo("s")
    for k, g in enumerate(graphs):
        n, e = g["n"], g["edge_index"].shape[1]
        assert torch.equal(sb.edge_index[:, e_off:e_off + e].cpu() - n_off, g["edge_index"])
        nt, ns = g["edge_index_t"].shape[1], g["edge_index_s"].shape[1]
        assert torch.equal(ei_t[:, t_off:t_off + nt].cpu() - n_off, g["edge_index_t"]), k
        assert torch.equal(ei_s[:, s_off:s_off + ns].cpu() - e_off, g["edge_index_s"]), k
        wt, ws = ew_t[t_off:t_off + nt].cpu(), ew_s[s_off:s_off + ns].cpu()
        assert torch.equal(torch.sign(wt), torch.sign(g["edge_weight_t"]))
        assert torch.equal(torch.sign(ws), torch.sign(g["edge_weight_s"]))
        assert torch.allclose(wt, g["edge_weight_t"], rtol=rtol, atol=0)
        assert torch.allclose(ws, g["edge_weight_s"], rtol=rtol, atol=0)
        assert abs(float(sb.lambda_max[k]) - float(g["maxeig"])) <= rtol * float(g["maxeig"])
        assert int(sb.num_edge1[k]) == e
        n_off, e_off, t_off, s_off = n_off + n, e_off + e, t_off + nt, s_off + ns
    assert ei_t.shape[1] == t_off and ei_s.shape[1] == s_off


def test_construction_vs_golden_reference_batch():
    """All golden graphs (incl. one with an isolated node) collated into ONE batch."""
    graphs = load_golden("construct.pt")
    off, src, dst, attr = 0, [], [], []
    for g in graphs:
        src.append(g["ei_dir"][0] + off)
        dst.append(g["ei_dir"][1] + off)
        attr.append(torch.arange(g["ei_dir"].shape[1]) % 3 + 1)
        off += g["n"]
    sb = build_simplex_batch(torch.cat(src).to(DEV), torch.cat(dst).to(DEV),
                             torch.tensor([g["n"] for g in graphs]), edge_attr=torch.cat(attr).to(DEV))
    _check_against(graphs, sb)
    assert torch.equal(sb.edge_attr.cpu(), torch.cat([g["edge_attr"] for g in graphs]))
    # tiny graph of SURVEY.md section 4
    ei = torch.tensor([[0, 0, 1, 2, 1, 2, 2, 3, 3], [1, 2, 2, 3, 0, 0, 1, 2, 3]])       # both directions + a self loop
    tb = build_simplex_batch(ei[0].to(DEV), ei[1].to(DEV), torch.tensor([4]))
    assert tb.edge_index.cpu().tolist() == [[0, 0, 1, 2], [1, 2, 2, 3]]
    assert abs(float(tb.lambda_max[0]) - 4.0) < 1e-6
    it, wt = tb.coo("t")
    assert it.cpu().tolist() == [[0, 0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 3], [0, 1, 2, 0, 1, 2, 0, 1, 2, 3, 2, 3]]
    assert torch.allclose(wt.cpu(), torch.tensor([1, -.5, -.5, -.5, 1, -.5, -.5, -.5, 1.5, -.5, -.5, .5]), atol=1e-6)
    is_, ws = tb.coo("s")
    assert is_.cpu().tolist() == [[0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3], [0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3]]
    assert torch.allclose(ws.cpu(), torch.tensor([1, .5, -.5, .5, 1, .5, -.5, -.5, .5, 1, -.5, -.5, -.5, 1]), atol=1e-6)


@pytest.mark.parametrize("shape,batch", [("zinc", 64), ("peptides", 8), ("cifar", 6), ("tsp", 2)])
def test_construction_vs_oracle_config_shapes(shape, batch):
    rng = np.random.default_rng(7)
    n_lo, n_hi, kind, param, _, _ = SHAPES[shape]
    graphs, src, dst, counts, off = [], [], [], [], 0
    for _ in range(batch):
        n = int(rng.integers(n_lo, n_hi + 1))
        ei = torch.from_numpy(np.ascontiguousarray(_tree_plus_chords(rng, n, param) if kind == "tree" else _knn_graph(rng, n, param)))
        ei_dir = torch.cat([ei, ei.flip(0)], 1)[:, torch.from_numpy(rng.permutation(2 * ei.shape[1]))]
        o = O.build_simplex_graph(ei_dir, n)
        graphs.append(dict(n=n, edge_index=o.edge_index, edge_index_t=o.edge_index_t, edge_weight_t=o.edge_weight_t,
                           edge_index_s=o.edge_index_s, edge_weight_s=o.edge_weight_s, maxeig=o.maxeig))
        src.append(ei_dir[0] + off)
        dst.append(ei_dir[1] + off)
        counts.append(n)
        off += n
    sb = build_simplex_batch(torch.cat(src).to(DEV), torch.cat(dst).to(DEV), torch.tensor(counts))
    _check_against(graphs, sb, rtol=1e-6)
    # properties: rows of L0 sum to 0, both operators symmetric, nnz identities
    it, wt = sb.coo("t")
    assert float(torch.zeros(sb.num_nodes, device=DEV).index_add_(0, it[0], wt).abs().max()) < 1e-5
    deg = sb.incidence.degree()
    assert it.shape[1] == int((deg > 0).sum() + 2 * sb.num_edges)
    assert sb.coo("s")[0].shape[1] == int(sb.num_edges + (deg * (deg - 1)).sum())


def test_model_runs_on_constructed_batch_like_host_built_batch():
    """The operators built on the GPU drive the conv exactly like the COO built on the host."""
    from hlhgat_b200 import functional as F_hl, _native as N
    from hlhgat_b200.simplex import CsrOperator
    b = make_batch("zinc", 32, seed=9)
    ei = b.edge_index
    sb = build_simplex_batch(torch.cat([ei[0], ei[1]]).to(DEV), torch.cat([ei[1], ei[0]]).to(DEV), b.num_node1)
    x = torch.randn(b.x_s.shape[0], 64, device=DEV)
    host_op = CsrOperator(b.edge_index_s.to(DEV), b.edge_weight_s.to(DEV), b.x_s.shape[0])
    (a,) = F_hl.poly_basis_fwd(N.HL_LAGUERRE, 4, [host_op], [x], 64)
    (c,) = F_hl.poly_basis_fwd(N.HL_LAGUERRE, 4, [sb.op_s], [x], 64)
    assert torch.allclose(a, c, rtol=1e-5, atol=1e-6)
    assert torch.equal(host_op.fwd[1], sb.op_s.fwd[1]) and torch.equal(host_op.fwd[0], sb.op_s.fwd[0])


def test_dataset_process_mirror_on_golden_graphs():
    """lib.Hodge_Dataset.simplex_batch_from_graphs = Dataset.process (lib/Hodge_Dataset.py:447-477) for a list of
    raw graphs with LOCAL node ids; same golden vectors as above."""
    from hlhgat_b200.lib.Hodge_Dataset import simplex_batch_from_graphs
    graphs = load_golden("construct.pt")
    sb = simplex_batch_from_graphs([g["ei_dir"] for g in graphs], [g["n"] for g in graphs], device=DEV,
                                   edge_attrs=[torch.arange(g["ei_dir"].shape[1]) % 3 + 1 for g in graphs])
    _check_against(graphs, sb)
    assert torch.equal(sb.edge_attr.cpu(), torch.cat([g["edge_attr"] for g in graphs]))
