import torch
from ..data import Batch, Data


def _collate(samples):
    first = samples[0]
    if isinstance(first, Data):
        return Batch.from_data_list(samples)
    if isinstance(first, (list, tuple)):
        return [_collate(list(s)) for s in zip(*samples)]
    return torch.utils.data.default_collate(samples)


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=_collate, **kwargs)
