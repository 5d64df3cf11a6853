"""Eigenvector positional encodings on the GPU for a whole mini-batch (reference
lib/Hodge_Dataset.py:97-112 `eig_pe` and its callers in `Dataset.process`, :457-458 / :586-587 / :846-847)."""
import torch

from . import _native as N
from .simplex import CsrOperator, csr_from_coo

__all__ = ["eig_pe_batch", "eig_pe", "flip_signs"]


def eig_pe_batch(op, counts, k, return_all=False, max_sweeps=30):
    """Encodings of every graph of a batch.  `op`: the block-diagonal operator (a `CsrOperator`, e.g. `SimplexBatch.op_t`
    / `.op_s`, i.e. L0 / L1 scaled by 2 / lambda_max); `counts` [G]: rows per graph (nodes for L0, edges for L1).
    Returns `pe` [R, k-1] (eigenvectors 1 .. k-1 in ascending eigenvalue order, zero-padded for graphs with fewer than
    k rows, largest-magnitude component positive) and `evals` [R] (ascending per graph); with `return_all` also the
    list of full eigenvector matrices per graph."""
    L = N.lib()
    rowptr, colidx, vals = op.fwd
    dev = colidx.device
    counts = torch.as_tensor(counts, dtype=torch.int64, device=dev)
    G = int(counts.numel())
    R = int(op.nrows)
    seg = torch.zeros(G + 1, dtype=torch.int32, device=dev)
    seg[1:] = torch.cumsum(counts, 0)
    mat = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    mat[1:] = torch.cumsum(counts * counts, 0)
    total, max_n = (int(mat[-1]), int(counts.max())) if G else (0, 0)
    if G and int(seg[-1]) != R:
        raise N.HlError("eig_pe_batch: counts do not add up to the operator's rows")
    pe = torch.zeros((R, max(k - 1, 0)), dtype=torch.float32, device=dev)
    evals = torch.zeros(R, dtype=torch.float32, device=dev)
    vecs = torch.empty(total, dtype=torch.float32, device=dev) if return_all else None
    sweeps = torch.zeros(max(G, 1), dtype=torch.int32, device=dev)
    nb = L.hl_eig_pe_workspace(total)
    ws = torch.empty(max(int(nb), 1), dtype=torch.uint8, device=dev)
    N.check(L.hl_eig_pe(seg.data_ptr(), G, max_n, mat.data_ptr(), total, rowptr.data_ptr(), colidx.data_ptr(), vals.data_ptr(),
                        k, evals.data_ptr(), pe.data_ptr(), max(k - 1, 1), N.ptr(vecs), sweeps.data_ptr(), max_sweeps,
                        ws.data_ptr(), nb, N.stream_ptr()), "hl_eig_pe")
    if return_all:
        offs, ns = mat.tolist(), counts.tolist()
        return pe, evals, [vecs[offs[g]:offs[g + 1]].view(ns[g], ns[g]) for g in range(G)], sweeps[:G]
    return pe, evals


def eig_pe(L, k=9):
    """The reference's signature (lib/Hodge_Dataset.py:97): one dense symmetric matrix (CUDA tensor) -> eigenvectors
    1 .. k-1 in ascending eigenvalue order, [n, min(k, n) - 1]."""
    if not L.is_cuda:
        raise N.HlError("eig_pe runs on CUDA tensors only (no CPU fallback)")
    n = L.shape[0]
    idx = L.nonzero().t().contiguous()
    op = CsrOperator.from_csr(*csr_from_coo(idx[0], idx[1], L[idx[0], idx[1]].float(), n)[:3], n)
    pe, _ = eig_pe_batch(op, torch.tensor([n]), k)
    return pe[:, : max(min(k, n) - 1, 0)]


def flip_signs(x, lead, generator=None):
    """The augmentation of the datasets' `get` (lib/Hodge_Dataset.py:429-439): one random +-1 per encoding column
    (columns `lead` .. of the feature matrix), drawn on the host like the reference does."""
    flips = (-1 + 2 * torch.randint(0, 2, (x.shape[1] - lead,), generator=generator)).to(x.dtype)
    return x * torch.cat([torch.ones(lead, dtype=x.dtype), flips]).to(x.device)
