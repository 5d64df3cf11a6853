class F1Score:  # imported by the reference, unused on the hot path
    def __init__(self, *a, **k):
        raise NotImplementedError
from . import classification  # noqa: E402,F401
