"""Hot-path subset of the reference's lib/Hodge_Dataset.py: `adj2par1` (:169-191)."""
from .Hodge_Cheb_Conv import adj2par1, degree  # noqa: F401
