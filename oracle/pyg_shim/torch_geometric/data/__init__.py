"""Data / Batch: attribute store + block-diagonal collation:
keys containing 'index' concatenate along dim -1 and are offset by __inc__ (default: num_nodes);
everything else concatenates along dim 0; python numbers collate to a [B] tensor."""
import copy
import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        object.__setattr__(self, "_store", {})
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos).items():
            if v is not None:
                self._store[k] = v
        for k, v in kwargs.items():
            self._store[k] = v

    # attribute protocol -------------------------------------------------------------------
    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        if key in ("x", "edge_index", "edge_attr", "y", "pos"):
            return None
        raise AttributeError(key)

    def __setattr__(self, key, value):
        prop = getattr(type(self), key, None)
        if isinstance(prop, property) and prop.fset is not None:
            prop.fset(self, value)
        else:
            self._store[key] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store

    @property
    def keys(self):
        return [k for k, v in self._store.items() if v is not None and not k.startswith("_")]

    @property
    def num_nodes(self):
        if "_num_nodes" in self._store:
            return self._store["_num_nodes"]
        x = self._store.get("x")
        if x is not None:
            return x.size(0)
        ei = self._store.get("edge_index")
        if ei is not None and ei.numel() > 0:
            return int(ei.max()) + 1
        return None

    @num_nodes.setter
    def num_nodes(self, v):
        self._store["_num_nodes"] = v

    @property
    def num_edges(self):
        ei = self._store.get("edge_index")
        return 0 if ei is None else ei.size(1)

    def __cat_dim__(self, key, value, *args, **kwargs):
        return -1 if "index" in key else 0

    def __inc__(self, key, value, *args, **kwargs):
        return self.num_nodes if "index" in key else 0

    def to(self, device, *a, **k):
        for key, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[key] = v.to(device, *a, **k)
        return self

    def cpu(self):
        return self.to("cpu")

    def clone(self):
        return copy.deepcopy(self)

    def __repr__(self):
        items = []
        for k in self.keys:
            v = self._store[k]
            items.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v}")
        return f"{type(self).__name__}({', '.join(items)})"


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list, follow_batch=None, exclude_keys=None):
        first = data_list[0]
        batch = cls()
        object.__setattr__(batch, "_proto", type(first))
        keys = [k for k in first._store.keys() if first._store[k] is not None and k != "_num_nodes"]
        incs = {k: 0 for k in keys}
        cols = {k: [] for k in keys}
        sizes = []
        for d in data_list:
            for k in keys:
                v = d._store[k]
                if torch.is_tensor(v) and v.dim() > 0:
                    inc = d.__inc__(k, v)
                    if isinstance(inc, torch.Tensor):
                        inc = int(inc)
                    cols[k].append(v + incs[k] if incs[k] != 0 else v)
                    incs[k] += inc
                else:
                    cols[k].append(v)
            sizes.append(d.num_nodes)
        for k in keys:
            v0 = cols[k][0]
            if torch.is_tensor(v0) and v0.dim() > 0:
                batch._store[k] = torch.cat(cols[k], dim=first.__cat_dim__(k, v0))
            elif torch.is_tensor(v0):
                batch._store[k] = torch.stack(cols[k])
            elif isinstance(v0, (int, float)):
                batch._store[k] = torch.tensor(cols[k])
            else:
                batch._store[k] = cols[k]
        if all(s is not None for s in sizes):
            n = torch.tensor(sizes, dtype=torch.long)
            batch._store["batch"] = torch.repeat_interleave(torch.arange(len(sizes)), n)
            batch._store["ptr"] = torch.cat([n.new_zeros(1), n.cumsum(0)])
            batch._store["_num_nodes"] = int(n.sum())
        batch._store["_num_graphs"] = len(data_list)
        return batch

    @property
    def num_graphs(self):
        return self._store["_num_graphs"]


class Dataset(torch.utils.data.Dataset):
    """Just enough of the PyG Dataset protocol for the reference classes to be *defined*."""

    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root = root
        self.transform = transform

    def len(self):
        raise NotImplementedError

    def get(self, idx):
        raise NotImplementedError

    def __len__(self):
        return self.len()

    def __getitem__(self, idx):
        return self.get(idx)


class InMemoryDataset(Dataset):
    pass


def download_url(*a, **k):
    raise RuntimeError("no network in this image")


def extract_zip(*a, **k):
    raise RuntimeError("no network in this image")
