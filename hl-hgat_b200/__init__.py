"""hlhgat_b200: B200-native (sm_100a) hot path of HL-HGAT -- Hodge-Laplacian polynomial convolution,
node<->edge simplex transfer, attention gate / cluster pooling, readout -- behind the reference's
PyTorch module API.  See DESIGN.md."""
from . import _native  # noqa: F401
from ._native import HlError, LIB_PATH  # noqa: F401
from .build import build  # noqa: F401
from . import simplex, functional, lanes  # noqa: F401
from .lanes import enable_lanes  # noqa: F401
from .functional import enable_project_then_transfer  # noqa: F401
from .dense_stack import enable_dense_stack  # noqa: F401
from .lib.Hodge_Cheb_Conv import (HodgeLaguerreConv, HodgeChebConv, HodgeLaguerreFastConv,  # noqa: F401
                                  NodeEdgeInt, MSI, SAPool, HL_filter, NEConv, GraphBatchNorm,
                                  adj2par1, degree)

__version__ = "0.1.0"
