"""The BASELINE.json configurations as concrete training workloads: model class and
constructor arguments of the reference script, synthetic batch generator, loss, per-GPU batch size.
Used by bench.py and the tests; the CPU oracle counterparts are looked up by class NAME in bench.py
(nothing in this package imports `oracle/`)."""
from types import SimpleNamespace

import torch
import torch.nn.functional as Fnn

from .synthetic import make_batch, make_multilevel_batch, make_tsp_batch


def focal_loss(preds, labels, valid=None, n_valid=None, alpha=0.25, gamma=2):
    """lib/Loss_function.py:14-26 (FocalLoss on the MEAN BCE-with-logits, times 1e4).  `valid` / `n_valid`:
    optional row mask and device row count of a padded batch (ghost rows excluded from the mean)."""
    if valid is None:
        logpt = -Fnn.binary_cross_entropy_with_logits(preds, labels)
    else:
        bce = Fnn.binary_cross_entropy_with_logits(preds, labels, reduction="none")
        logpt = -(bce * valid).sum() / (n_valid.reshape(()).to(bce.dtype) * bce.shape[1])
    pt = torch.exp(logpt)
    return -((1 - pt) ** gamma) * alpha * logpt * 1e4


def _graph_level(loss):
    def fn(model, batch):
        b0 = batch[0] if not hasattr(batch, "y") else batch
        out = model(batch, device=b0.x_t.device)
        return loss(out[: b0.num_graphs], b0.y)
    return fn


def _attpool_loss(loss):
    def fn(model, batch):                       # the scripts call model(data, if_att=True) (main_cifar...py:135, main_pepfunc...py:181)
        out, _, _ = model(batch, device=batch[0].x_t.device, if_att=True)
        return loss(out[: batch[0].num_graphs], batch[0].y)
    return fn


def _tsp_loss(model, batch):
    out, _ = model(batch, device=batch.x_t.device)
    y = batch.y.view(-1, 1)
    nv = getattr(batch, "n_valid_edges", None)
    if nv is None:
        return focal_loss(out, y)
    valid = (torch.arange(out.shape[0], device=out.device) < nv).to(out.dtype).view(-1, 1)
    return focal_loss(out, y, valid, nv)


def _zinc_batch(batch, seed):
    return make_batch("zinc", batch, seed=seed)


def _pep_batch(batch, seed):
    d = make_multilevel_batch("peptides", batch, seed=seed, node_dim=19, edge_dim=13, num_targets=10)
    d[0].y = (d[0].y > 0.8).float()                       # multi-label targets (peptides-func: 10 classes)
    return d


def _cifar_batch(batch, seed):
    d = make_multilevel_batch("cifar", batch, seed=seed, node_dim=15, edge_dim=14, num_targets=1)
    d[0].y = torch.randint(0, 10, (batch,), generator=torch.Generator().manual_seed(seed))
    return d


def _tsp_batch(batch, seed):
    b = make_tsp_batch(batch, seed=seed)
    b.y = b.y.float()
    return b


WORKLOADS = {
    # BASELINE.json configs[1] (configs[0] is the same model at batch 128 on the CPU)
    "zinc": SimpleNamespace(
        model="HL_HGCNN_zinc_dense_int3_pyr",
        ctor=dict(channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[], K=2, node_dim=21, edge_dim=3, keig=7),
        batch=1024, cpu_sample=1024, make=_zinc_batch, levels=1, deg_eps=0.0, long_rows=False,
        loss=_graph_level(torch.nn.L1Loss()), label="zinc_pyr_train_b1024_K2_fp32",
        metric="train graphs/sec ZINC-shaped (HL_HGCNN_zinc_dense_int3_pyr, batch 1024/GPU, K=2, fp32)"),
    # the same model with the defaults of main_zinc_HL_HGCNN_dense_int3_pyr.py:26-35 (SURVEY section 8d, config 2 "also report")
    "zinc_default": SimpleNamespace(
        model="HL_HGCNN_zinc_dense_int3_pyr",
        ctor=dict(channels=[2, 3, 3], filters=[64, 128, 256], mlp_channels=[256, 256], K=6, node_dim=21, edge_dim=3, keig=7),
        batch=1024, cpu_sample=128, make=_zinc_batch, levels=1, deg_eps=0.0, long_rows=False,
        loss=_graph_level(torch.nn.L1Loss()), label="zinc_pyr_train_b1024_K6_c233_mlp256x2_fp32",
        metric="train graphs/sec ZINC-shaped (HL_HGCNN_zinc_dense_int3_pyr, script defaults: channels [2,3,3], K=6, mlp [256,256]; batch 1024/GPU, fp32)"),
    # configs[2]: main_pepfunc_HL_HGCNN_dense_int3_attpool.py:249-254 (script defaults)
    "peptides": SimpleNamespace(
        model="HL_HGCNN_pepfunc_dense_int3_attpool",
        ctor=dict(channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[256], pool_loc=1, K=6, node_dim=9, edge_dim=3,
                  keig=10, num_classes=10),
        batch=64, cpu_sample=16, make=_pep_batch, levels=2, deg_eps=1e-6, long_rows=False,
        loss=_attpool_loss(focal_loss), label="pepfunc_attpool_train_b64_K6_fp32",
        metric="train graphs/sec peptides-func-shaped (HL_HGCNN_pepfunc_dense_int3_attpool, batch 64/GPU, K=6, fp32)"),
    # configs[3]: lib/Hodge_ST_Model.py:958, main_cifar10SP_HL_HGCNN_dense_int3_attpool.py:36
    "cifar": SimpleNamespace(
        model="HL_HGCNN_CIFAR10SP_dense_int3_attpool",
        ctor=dict(channels=[2, 2, 2], filters=[64, 128, 256], mlp_channels=[256], K=4, node_dim=5, edge_dim=4, keig=10,
                  pool_loc=1, l=0.5, num_classes=10),
        batch=256, cpu_sample=16, make=_cifar_batch, levels=2, deg_eps=1e-6, long_rows=True,
        loss=_attpool_loss(torch.nn.CrossEntropyLoss()), label="cifar10sp_attpool_train_b256_K4_fp32",
        metric="train graphs/sec CIFAR10-superpixel-shaped (HL_HGCNN_CIFAR10SP_dense_int3_attpool, batch 256/GPU, K=4, fp32)"),
    # configs[4]: lib/Hodge_ST_Model.py:756, main_TSP_HL_HGCNN_dense_int3_pyr.py:38-48
    "tsp": SimpleNamespace(
        model="HL_HGCNN_TSP_dense_int3_pyr",
        ctor=dict(channels=[4, 4, 4], filters=[32, 64, 128], mlp_channels=[256], K=4, node_dim=2, edge_dim=1, num_classes=1),
        batch=32, cpu_sample=1, make=_tsp_batch, levels=1, deg_eps=1e-6, long_rows=True,
        loss=_tsp_loss, label="tsp_pyr_train_b32_K4_fp32",
        metric="train graphs/sec TSP-shaped (HL_HGCNN_TSP_dense_int3_pyr, 500-node kNN-25 graphs, batch 32/GPU, K=4, fp32)"),
}
