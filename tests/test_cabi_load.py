"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/hlhgat.h declares, and the product path refuses CPU tensors instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

import hlhgat_b200 as H
from hlhgat_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hlhgat.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hl_[a-z0-9_]+)\s*\(", src)))


def test_library_built_and_exports_every_declared_symbol():
    H.build()
    handle = ctypes.CDLL(H.LIB_PATH)
    names = _declared()
    assert len(names) >= 17
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/hlhgat.h but not exported"
    assert sorted(N.exported_symbols()) == names, "ctypes binding and header drifted apart"
    assert N.lib().hl_version() == 1
    assert N.lib().hl_status_string(0) == b"ok"


def test_no_cpu_fallback():
    conv = H.HodgeLaguerreConv(4, 5, 3)
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(H.HlError):
        conv(torch.randn(2, 4), ei, torch.ones(2))
    with pytest.raises(H.HlError):
        H.functional.bn_act_train(torch.randn(4, 4), torch.ones(4), torch.zeros(4))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", "/nonexistent/libhlhgat.so")
    with pytest.raises(H.HlError, match="no CPU or PyTorch fallback"):
        N.lib()


def test_state_dict_contract_matches_reference_names():
    """Key names / shapes of SURVEY.md section 4 (checkpoint contract)."""
    from hlhgat_b200.lib.Hodge_ST_Model import HL_HGCNN_zinc_dense_int3_pyr
    from conftest import load_golden
    z = load_golden("zinc_model.pt")
    m = HL_HGCNN_zinc_dense_int3_pyr(K=2, **z["ctor"])
    m.load_state_dict(z["runs"][2]["state"], strict=True)
    keys = set(m.state_dict())
    assert "HL_init_conv.module_0.lins.1.weight" in keys and "HL_init_conv.module_1.module.running_var" in keys
    assert "NEInt00.WV_Node.0.weight" in keys and "NEConv10.module_4.bias" in keys and "out.weight" in keys
    conv = H.HodgeLaguerreConv(3, 7, 4)
    assert [tuple(l.weight.shape) for l in conv.lins] == [(7, 3)] * 4 and conv.bias.abs().sum() == 0
    with pytest.raises(AssertionError):
        H.HodgeLaguerreConv(3, 7, 0)


def test_pairdata_collate_matches_reference_batch():
    """hlhgat_b200.lib.Hodge_Dataset.PairData + collate reproduce the batch PyG's DataLoader built from the
    reference's PairData.__inc__ (golden: tests/golden/zinc_model.pt), and the list-of-levels form."""
    import torch
    from conftest import load_golden
    from oracle import hodge_oracle as O
    from hlhgat_b200.lib.Hodge_Dataset import PairData, collate
    z = load_golden("zinc_model.pt")
    b = z["batch"]
    graphs, off_n, off_e = [], 0, 0
    for g in z["raw"]:
        c = O.build_simplex_graph(g["ei_dir"], g["n"])
        e = c.edge_index.shape[1]
        d = PairData(x_s=b["x_s"][off_e:off_e + e], edge_index_s=c.edge_index_s, edge_weight_s=c.edge_weight_s,
                     x_t=b["x_t"][off_n:off_n + g["n"]], edge_index_t=c.edge_index_t, edge_weight_t=c.edge_weight_t,
                     y=torch.zeros(1))
        d.num_node1, d.num_edge1, d.num_nodes, d.edge_index = g["n"], e, g["n"], c.edge_index
        graphs.append(d)
        off_n, off_e = off_n + g["n"], off_e + e
    bb = collate(graphs)
    for k in ("x_t", "x_s", "edge_index", "edge_index_t", "edge_index_s", "edge_weight_t", "edge_weight_s", "num_node1", "num_edge1"):
        assert torch.equal(getattr(bb, k), b[k]), k
    assert bb.num_graphs == len(graphs) and bb.num_nodes == off_n and bb.ptr[-1] == off_n
    levels = collate([[g, g] for g in graphs])
    assert isinstance(levels, list) and len(levels) == 2 and torch.equal(levels[1].edge_index_s, b["edge_index_s"])


def test_lanes_are_inert_without_cuda_tensors():
    """lanes.open_lanes yields None (single stream) unless enabled AND on a CUDA device; join is a no-op then."""
    from hlhgat_b200 import lanes
    assert not lanes.lanes_enabled() and lanes.active() is None
    with lanes.open_lanes("cpu") as ln:
        assert ln is None
    lanes.enable_lanes(True)
    try:
        with lanes.open_lanes("cpu") as ln:
            assert ln is None and lanes.active() is None
    finally:
        lanes.enable_lanes(False)
    lanes.join()


def test_split_desc_layout_matches_header():
    """hl_split_desc (include/hlhgat.h) as bound by ctypes: 3 pointers, 2 int64, 4 int32 = 56 bytes at their natural
    offsets -- functional.WeightSplitPlan uploads an array of these to the device as raw bytes."""
    import ctypes as C
    import re
    from hlhgat_b200 import _native as N
    d = N.SplitDesc
    offs = {f: getattr(d, f).offset for f, _ in d._fields_}
    assert offs == {"src": 0, "hi": 8, "lo": 16, "ld_src": 24, "ld_out": 32, "rows": 40, "cols": 44, "transpose": 48, "reserved": 52}
    assert C.sizeof(d) == 56
    header = open(os.path.join(ROOT, "include", "hlhgat.h")).read()
    body = re.search(r"typedef struct hl_split_desc \{(.*?)\} hl_split_desc;", header, re.S).group(1)
    names = re.findall(r"(\w+);", body)
    assert names == [f for f, _ in d._fields_]           # same members in the same order as the C declaration


def test_wgrad_reduce_desc_layout_matches_header():
    """hl_wgrad_reduce_desc as bound by ctypes: 5 pointers, 3 int64, 8 int32 = 96 bytes, same member order as the header
    (functional.WgradReducePlan hands a host array of these to hl_wgrad_reduce_batch)."""
    import ctypes as C
    import re
    from hlhgat_b200 import _native as N
    d = N.WgradReduceDesc
    assert C.sizeof(d) == 96
    offs = [getattr(d, f).offset for f, _ in d._fields_]
    assert offs == [0, 8, 16, 24, 32, 40, 48, 56, 64, 68, 72, 76, 80, 84, 88, 92]
    header = open(os.path.join(ROOT, "include", "hlhgat.h")).read()
    body = re.search(r"typedef struct hl_wgrad_reduce_desc \{(.*?)\} hl_wgrad_reduce_desc;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"(\w+);", body)
    assert names == [f for f, _ in d._fields_]


def test_wgrad_reduce_plan_overlap_rule():
    """Destinations of one batch: column blocks of one matrix do not conflict, the same view or a straddling one does."""
    import torch
    from hlhgat_b200.functional import WgradReducePlan
    gw = torch.zeros(8, 12)
    plan = WgradReducePlan()
    assert plan.claim(gw[:, :4], gw[:, 4:8])
    assert plan.claim(gw[:, 8:])
    assert not plan.claim(gw[:, 2:6])
    assert not plan.claim(gw)
    assert not plan.claim(gw[:, :4])
    assert plan.claim(torch.zeros(8), torch.zeros(3, 3))


def test_bn_epilogue_request_is_served_once_and_scoped():
    """functional.bn_stats_from_epilogue: the request is visible to the first taker inside the `with`, gone afterwards,
    restored on exit, and absent when disabled; a BnTiles without statistics never matches."""
    import torch
    from hlhgat_b200 import functional as F
    assert F._take_bn_request() is None
    with F.bn_stats_from_epilogue(None) as req:
        assert (req is not None) == F._BN_EPILOGUE
        first = F._take_bn_request()
        assert first is req and F._take_bn_request() is None
        with F.bn_stats_from_epilogue(None, enabled=False) as inner:
            assert inner is None and F._take_bn_request() is None
    assert F._take_bn_request() is None
    if req is not None:
        assert not req.matches(torch.zeros(4, 4), None)
