"""torch-sparse 0.6.15 subset: SparseTensor(row, col, value).t() and matmul(adj, x, 'add')."""
import torch


class SparseTensor:
    def __init__(self, row=None, col=None, value=None, sparse_sizes=None):
        n = int(max(row.max(), col.max())) + 1 if sparse_sizes is None else None
        self.sizes = sparse_sizes if sparse_sizes is not None else (n, n)
        key = row * self.sizes[1] + col
        perm = torch.argsort(key, stable=True)
        self.row, self.col = row[perm], col[perm]
        self.value = value[perm] if value is not None else None

    def t(self):
        return SparseTensor(self.col, self.row, self.value, (self.sizes[1], self.sizes[0]))

    def to(self, *a, **k):
        return self


def matmul(src, other, reduce="sum"):
    assert reduce in ("sum", "add")
    v = src.value if src.value is not None else torch.ones(src.row.numel(), dtype=other.dtype)
    msg = v.view(-1, 1) * other.index_select(0, src.col)
    out = torch.zeros((src.sizes[0],) + tuple(other.shape[1:]), dtype=other.dtype)
    return out.index_add_(0, src.row, msg)
