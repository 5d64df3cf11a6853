"""CPU stand-in for the subset of torch_geometric the reference imports (test infrastructure)."""
__version__ = "0.0-shim"
from . import typing, utils, data, nn, loader, datasets  # noqa: F401
