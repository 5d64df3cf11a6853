"""the reference's cached per-graph `.pt` dicts ({'graph': PairData | [PairData...], 'maxeig',
'par1'}, lib/Hodge_Dataset.py:475-476, :528-529) read WITHOUT torch_geometric and fed to the B200 path.  The fixtures under
tests/golden/cache/ were written by tests/golden/make_cache_fixtures.py: graph content from the unmodified reference code,
pickled under the real class paths in both PyG layouts (>= 2.0 `_store` / `_mapping`, 1.x plain `__dict__`)."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

from conftest import GOLDEN
from hlhgat_b200.lib import cached_graphs as CG

CACHE = os.path.join(GOLDEN, "cache")
KEYS = ("x_t", "x_s", "edge_index", "edge_index_t", "edge_index_s", "edge_weight_t", "edge_weight_s", "y")


def _same(g, exp):
    for k in KEYS:
        if exp.get(k) is None:                           # coarse levels carry no target
            assert getattr(g, k) is None, k
            continue
        assert torch.equal(getattr(g, k), exp[k]), k
    assert int(g.num_node1) == int(exp["num_node1"]) and int(g.num_edge1) == int(exp["num_edge1"])


def test_reader_decodes_both_pyg_layouts_without_torch_geometric():
    exp = torch.load(os.path.join(CACHE, "expected.pt"), weights_only=False)
    for i in range(4):                                   # even files: PyG >= 2.0 layout, odd files: PyG 1.x layout
        g, extra = CG.load_cached_graph(os.path.join(CACHE, f"ZINC_BM_alleig_{i + 1}.pt"))
        _same(g, exp["single"][i])
        assert float(extra["maxeig"]) == float(exp["maxeig"][i])
        n, e = g.x_t.shape[0], g.x_s.shape[0]
        assert tuple(extra["par1"].shape) == (n, e)      # the dense boundary matrix the reference also stores
        assert torch.equal(extra["par1"].abs().sum(0), torch.full((e,), 2.0))
    for i in range(3):
        levels, _ = CG.load_cached_graph(os.path.join(CACHE, f"ZINC_BM_MLGC_{i + 1}.pt"))
        assert isinstance(levels, list) and len(levels) == 2
        for lv, e in zip(levels, exp["multi"][i]):
            _same(lv, e)
    # nothing of torch_geometric was imported to do so (unless another test put the shim there)
    mod = sys.modules.get("torch_geometric")
    assert mod is None or "pyg_shim" in (getattr(mod, "__file__", "") or "")


def test_reader_refuses_foreign_code():
    import pickle
    import io

    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned > /tmp/hl_pwned",))

    buf = io.BytesIO()
    torch.save({"graph": Evil()}, buf)
    buf.seek(0)
    path = os.path.join("/tmp", "hl_evil.pt")
    open(path, "wb").write(buf.getvalue())
    if os.path.exists("/tmp/hl_pwned"):
        os.remove("/tmp/hl_pwned")
    with pytest.raises(Exception):
        CG.load_cached_graph(path)
    assert not os.path.exists("/tmp/hl_pwned")
    assert pickle is not None


def test_get_pads_truncates_and_flips_like_the_reference():
    """lib/Hodge_Dataset.py:425-440: keep node_dim + keig - 1 columns, zero-pad small graphs, one random sign per kept
    eigenvector column (applied only to truncated samples in the single-level class)."""
    exp = torch.load(os.path.join(CACHE, "expected.pt"), weights_only=False)
    for keig in (4, 40):
        ds = CG.ZINC_HG_BM_par1_EigPE(CACHE, keig=keig, generator=torch.Generator().manual_seed(5))
        assert len(ds) == 4
        gen = torch.Generator().manual_seed(5)
        for i in range(4):
            g = ds.get(i)
            for name, raw in (("x_t", 21), ("x_s", 3)):
                full = exp["single"][i][name]
                width = raw + keig - 1
                got = getattr(g, name)
                assert got.shape == (full.shape[0], width)
                if full.shape[1] < width:                                     # padded, not flipped
                    assert torch.equal(got[:, :full.shape[1]], full) and float(got[:, full.shape[1]:].abs().max()) == 0.0
                else:
                    sign = torch.cat([torch.ones(raw), (-1 + 2 * torch.randint(0, 2, (keig - 1,), generator=gen)).float()])
                    assert torch.equal(got, full[:, :width] * sign)
    ml = CG.ZINC_HG_BM_par1_MLGC(CACHE, keig=4, sign_flip=False)
    lv = ml.get(1)
    assert lv[0].x_t.shape[1] == 21 + 1 + 3 and lv[0].x_s.shape[1] == 3 + 1 + 3       # cluster-id column kept in front
    assert torch.equal(lv[0].x_t[:, 0], exp["multi"][1][0]["x_t"][:, 0])


def test_collated_cache_batch_matches_oracle_collate():
    from oracle import hodge_oracle as O
    ds = CG.ZINC_HG_BM_par1_EigPE(CACHE, keig=4, sign_flip=False)
    b = ds.batch(range(4))
    ref = O.collate([SimpleNamespace(**{k: getattr(ds.get(i), k) for k in KEYS}) for i in range(4)])
    for k in KEYS:
        assert torch.equal(getattr(b, k).reshape(-1) if k == "y" else getattr(b, k), getattr(ref, k).reshape(-1) if k == "y" else getattr(ref, k)), k
    assert b.num_graphs == 4 and torch.equal(torch.as_tensor(b.num_node1), torch.as_tensor(ref.num_node1))


@pytest.mark.gpu
def test_cached_graphs_feed_the_b200_model():
    """cache files -> get -> collate -> device -> HL_HGCNN_zinc_dense_int3_pyr on the GPU == the oracle on the same batch."""
    from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
    from oracle import hodge_oracle as O
    torch.manual_seed(0)
    keig = 5
    ds = CG.ZINC_HG_BM_par1_EigPE(CACHE, keig=keig, sign_flip=False)
    ctor = dict(channels=[1, 1], filters=[16, 32], mlp_channels=[], K=3, node_dim=21, edge_dim=3, keig=keig - 1)
    ref = O.HL_HGCNN_zinc_dense_int3_pyr(**ctor).train()
    model = HL_HGCNN_zinc_dense_int3_pyr(**ctor).to("cuda:0").train()
    model.load_state_dict(ref.state_dict(), strict=True)
    host = ds.batch(range(4))
    want = ref(host)
    dev = ds.batch(range(4), device="cuda:0")
    got = model(dev, device="cuda:0")
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-5), (got.cpu() - want).abs().max()
    # two-level samples -> the attention-pooling caller
    from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_pepfunc_dense_int3_attpool
    ml = CG.ZINC_HG_BM_par1_MLGC(CACHE, keig=keig, sign_flip=False)
    c2 = dict(channels=[1, 1], filters=[16, 32], mlp_channels=[16], K=2, node_dim=21, edge_dim=3, keig=keig - 1, pool_loc=0,
              num_classes=3)
    ref2 = O.HL_HGCNN_pepfunc_dense_int3_attpool(**c2).train()
    m2 = HL_HGCNN_pepfunc_dense_int3_attpool(**c2).to("cuda:0").train()
    m2.load_state_dict(ref2.state_dict(), strict=True)
    want2 = ref2(ml.batch(range(3)))
    got2 = m2(ml.batch(range(3), device="cuda:0"), device="cuda:0")
    assert torch.allclose(got2.cpu(), want2, rtol=1e-4, atol=1e-5), (got2.cpu() - want2).abs().max()
