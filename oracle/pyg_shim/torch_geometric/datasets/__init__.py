class _NoNetwork:
    def __init__(self, *a, **k):
        raise RuntimeError("dataset download needs network; use the synthetic generators")


class GNNBenchmarkDataset(_NoNetwork):
    pass


class ZINC(_NoNetwork):
    pass
