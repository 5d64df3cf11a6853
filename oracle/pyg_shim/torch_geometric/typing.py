from typing import Optional
from torch import Tensor

OptTensor = Optional[Tensor]
try:  # DEMO imports SparseTensor from here
    from torch_sparse import SparseTensor  # noqa: F401
except Exception:  # pragma: no cover
    SparseTensor = None
