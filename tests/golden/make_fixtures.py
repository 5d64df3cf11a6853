"""Fixtures derived from the artefacts the reference ships (build container only: reads /root/reference).

  python tests/golden/make_fixtures.py

* brain_ckpt_contract.json -- state_dict keys and shapes of HL-HGAT-DEMO/weights/HL_HGAT_Brain.pt (the only checkpoint
  in the reference): the naming / shape contract our modules must satisfy (module_{i} children of gnn.Sequential,
  `lins.{k}.weight [out,in]`, gnn.BatchNorm's `.module`, NodeEdgeInt's WV_* / WQ_* / WK_*).
* group_fc.pt -- the real brain skeleton of the DEMO (notebook cell 46: HL-HGAT-DEMO/data/Group_FC.mat masked by
  Group_FCMask.mat, upper triangle): undirected edge list + edge values, and the known answers of the reference's own
  construction formulas on it (lib/Hodge_Dataset.py:451-456,467-468 evaluated densely here, fp32): lambda_max, nnz and
  checksums of L0 / L1, degrees, sampled entries, and the low end of the L0 spectrum (for the eigenvector encodings).
"""
import json
import os
import sys

import numpy as np
import torch
from scipy.io import loadmat

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
sys.path.insert(0, "/root/reference")
DEMO = "/root/reference/HL-HGAT-DEMO"

import lib.Hodge_Dataset as RD  # noqa: E402
from torch_geometric.utils import dense_to_sparse  # noqa: E402  (shim)


def brain_contract():
    sd = torch.load(os.path.join(DEMO, "weights", "HL_HGAT_Brain.pt"), map_location="cpu", weights_only=False)
    out = {k: list(v.shape) for k, v in sd.items()}
    with open(os.path.join(HERE, "brain_ckpt_contract.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("brain_ckpt_contract.json", len(out), "keys")


def group_fc():
    torch.set_num_threads(1)
    fc = torch.tensor(loadmat(os.path.join(DEMO, "data", "Group_FC.mat"))["fc_mean"])
    fc[fc < 0] = 0.001
    mask = torch.tensor(loadmat(os.path.join(DEMO, "data", "Group_FCMask.mat"))["sf_mask"])
    skeleton = torch.triu(fc * mask, diagonal=1).to_sparse()                 # notebook cell 46
    ei = skeleton.indices()
    n = int(ei.max()) + 1
    e = ei.shape[1]
    par1 = RD.adj2par1(ei, n, e).to_dense()
    L0 = torch.matmul(par1, par1.T)
    lam, _ = torch.linalg.eigh(L0)
    maxeig = lam.max()
    L0 = 2 * torch.matmul(par1, par1.T) / maxeig
    L1 = 2 * torch.matmul(par1.T, par1) / maxeig
    eit, ewt = dense_to_sparse(L0)
    eis, ews = dense_to_sparse(L1)
    deg = torch.zeros(n).index_add_(0, ei.reshape(-1), torch.ones(2 * e))
    pick = torch.linspace(0, eis.shape[1] - 1, 4096).long()
    ev0 = torch.linalg.eigvalsh(L0.double())
    out = dict(edge_index=ei.to(torch.int16), edge_value=skeleton.values().float(), num_nodes=n, num_edges=e,
               maxeig=maxeig.clone(), maxeig64=torch.linalg.eigvalsh((par1 @ par1.T).double()).max(),
               nnz_t=int(eit.shape[1]), nnz_s=int(eis.shape[1]), max_degree=int(deg.max()),
               sum_w_t=float(ewt.double().sum()), sum_abs_w_s=float(ews.double().abs().sum()), sum_w_s=float(ews.double().sum()),
               pick=pick, pick_ei_s=eis[:, pick].to(torch.int32), pick_w_s=ews[pick].clone(),
               ei_t=eit.to(torch.int16), w_t=ewt.clone(), l0_spectrum_low=ev0[:24].clone(), l0_spectrum_high=ev0[-8:].clone())
    path = os.path.join(HERE, "group_fc.pt")
    torch.save(out, path)
    print("group_fc.pt", os.path.getsize(path), "bytes; N", n, "E", e, "maxeig", float(maxeig), "nnz", eit.shape[1], eis.shape[1],
          "max degree", int(deg.max()))


if __name__ == "__main__":
    brain_contract()
    group_fc()
