"""Drop-in mirror of the reference's `lib` package for the hot path: put `hl-hgat_b200/` ahead of
the reference tree on sys.path and `from lib.Hodge_Cheb_Conv import *` resolves here."""
