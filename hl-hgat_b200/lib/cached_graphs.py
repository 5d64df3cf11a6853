"""PyG-free reader of the reference's on-disk graph caches.

`Dataset.process` of the reference writes one file per graph, `torch.save({'graph': PairData | [PairData, ...],
'maxeig': ..., 'par1': ...}, '<NAME>_<i+1>.pt')` (lib/Hodge_Dataset.py:475-476, :528-529, :603-604, :663-664, :744),
and `get(idx)` reads it back, pads / truncates the eigenvector encodings to `keig - 1` columns and flips their signs at
random (:425-440, :500-515, :554-569, :628-650).  Those files pickle `lib.Hodge_Dataset.PairData`, a subclass of
`torch_geometric.data.Data`, so the reference can only read them with torch_geometric importable.  Here the pickle
stream is decoded WITHOUT importing either: every class that is not a torch tensor / storage / plain container is
materialised as an inert attribute bag, from which the tensors are lifted into `hlhgat_b200.lib.Hodge_Dataset.PairData`.
Both PyG layouts are understood -- PyG >= 2.0 (`Data.__dict__ = {'_store': GlobalStorage}` whose state holds
`_mapping`) and PyG 1.x (the attributes directly in `__dict__`).  No foreign code runs: `find_class` never imports a
module outside an allow-list (torch's tensor rebuild helpers / storage classes / dtypes, collections.OrderedDict, numpy
array reconstruction) -- anything else in the stream becomes an inert stand-in object.
"""
import glob
import os
import pickle
import re
import types

import torch

from .Hodge_Dataset import PairData, collate

__all__ = ["load_cached_graph", "CachedGraphs", "ZINC_HG_BM_par1_EigPE", "ZINC_HG_BM_par1_MLGC", "Peptides_Func_EigPE",
           "Peptides_Func_EigPE_MLGC", "TSP_EigPE", "collate"]

_BUILTINS_OK = {"dict", "list", "tuple", "set", "frozenset", "int", "float", "bool", "str", "bytes", "bytearray", "complex",
                "slice", "range", "object"}
_EXACT_OK = {("collections", "OrderedDict"), ("_codecs", "encode"), ("copyreg", "_reconstructor"),
             ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"),
             ("numpy._core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar")}


def _torch_object_ok(obj, name):
    """What a tensor pickle needs from `torch`: the rebuild helpers, storage classes, dtypes, Size / device -- never an
    arbitrary callable of the package (torch.load, torch.hub, ... stay unreachable)."""
    if isinstance(obj, (torch.dtype, torch.layout, torch.memory_format)) or obj in (torch.Size, torch.device, torch.Tensor):
        return True
    if isinstance(obj, type) and (name.endswith("Storage") or issubclass(obj, torch.Tensor)):
        return True
    return callable(obj) and name.startswith("_rebuild")


class _Bag:
    """Stand-in for an unpicklable foreign class: keeps the pickled state, runs nothing."""
    _hl_origin = ("?", "?")

    def __init__(self, *args, **kwargs):
        self.__dict__["_hl_args"] = (args, kwargs)

    def __setstate__(self, state):
        self.__dict__["_hl_state"] = state

    def __setitem__(self, k, v):                       # dict-like pickles (MutableMapping subclasses saved item by item)
        self.__dict__.setdefault("_hl_items", {})[k] = v

    def __reduce__(self):                              # never re-pickled as the foreign class
        raise pickle.PicklingError("cached-graph stand-ins are read-only")


_BAGS = {}


def _bag_class(module, name):
    key = (module, name)
    cls = _BAGS.get(key)
    if cls is None:
        cls = _BAGS[key] = type(name, (_Bag,), {"_hl_origin": key})
    return cls


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        root = module.split(".")[0]
        if (module, name) in _EXACT_OK or (module == "builtins" and name in _BUILTINS_OK):
            return super().find_class(module, name)
        if root == "torch":
            obj = super().find_class(module, name)
            if _torch_object_ok(obj, name):
                return obj
        return _bag_class(module, name)                # torch_geometric.*, lib.Hodge_Dataset.PairData, anything else


def _pickle_module():
    m = types.ModuleType("hlhgat_b200_cached_graph_pickle")
    m.__dict__.update({k: getattr(pickle, k) for k in ("load", "loads", "dump", "dumps", "HIGHEST_PROTOCOL", "PickleError",
                                                       "UnpicklingError", "Pickler")})
    m.Unpickler = _Unpickler
    return m


def _attributes(obj):
    """The attribute mapping of a pickled `Data`-like object, whichever PyG version wrote it."""
    if isinstance(obj, dict):
        return obj
    if not isinstance(obj, _Bag):
        raise TypeError(f"cannot interpret {type(obj)} as a graph")
    state = obj.__dict__.get("_hl_state")
    if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):     # (None, slots-state) convention
        state = {**(state[0] or {}), **state[1]}
    if state is None:
        state = obj.__dict__.get("_hl_items", {})
    if "_store" in state:                               # PyG >= 2.0: Data.__dict__['_store'] = GlobalStorage(_mapping=...)
        inner = _attributes(state["_store"])
        return inner.get("_mapping", inner)
    if "_mapping" in state:
        return state["_mapping"]
    return state                                        # PyG 1.x: the attributes themselves


_KEYS = ("edge_index_s", "x_s", "edge_index_t", "x_t", "edge_weight_s", "edge_weight_t", "edge_index", "y")


def _to_pairdata(obj):
    attrs = _attributes(obj)
    if not any(torch.is_tensor(attrs.get(k)) for k in _KEYS):
        raise ValueError(f"no graph tensors in pickled object of class {'.'.join(getattr(obj, '_hl_origin', ('?',)))}")
    g = PairData(**{k: attrs.get(k) for k in _KEYS})
    for k, v in attrs.items():
        if k in _KEYS or k.startswith("_") or isinstance(v, _Bag):
            continue
        setattr(g, k, v)                                # num_node1, num_edge1, num_nodes, pos_t, pos_s, ...
    if getattr(g, "num_node1", None) is None and g.x_t is not None:
        g.num_node1 = int(g.x_t.shape[0])
    if getattr(g, "num_edge1", None) is None and g.x_s is not None:
        g.num_edge1 = int(g.x_s.shape[0])
    return g


def load_cached_graph(path, map_location="cpu"):
    """One cache file of the reference -> `PairData`, or a list of `PairData` (fine graph + coarsened levels).  Also
    returns the file's other entries (`maxeig`, `par1`) as a dict."""
    blob = torch.load(path, map_location=map_location, pickle_module=_pickle_module(), weights_only=False)
    if not isinstance(blob, dict) or "graph" not in blob:
        raise ValueError(f"{path}: not a cache file of the reference ({{'graph': ...}} expected)")
    graph = blob["graph"]
    graph = [_to_pairdata(g) for g in graph] if isinstance(graph, (list, tuple)) else _to_pairdata(graph)
    extra = {k: v for k, v in blob.items() if k != "graph" and not isinstance(v, _Bag)}
    return graph, extra


def _natural(path):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", os.path.basename(path))]


class CachedGraphs:
    """`get(idx)` of the reference's cached datasets over a directory of `<prefix><i+1>.pt` files.

    node_dim / edge_dim: the raw feature widths in front of the eigenvector encodings; keig: the model's `keig`
    (keig - 1 encoding columns are kept, zero-padded for small graphs, :428-439); `levels` > 1: files hold a list
    `[fine, coarse, ...]` whose fine features carry the cluster id in column 0 (:503-514); sign_flip: the random
    +-1 per encoding column of every `get` (:429,:434); `generator` makes it reproducible."""

    def __init__(self, root, prefix, node_dim, edge_dim, keig, levels=1, sign_flip=True, generator=None, crop=None):
        self.files = sorted(glob.glob(os.path.join(root, prefix + "*.pt")), key=_natural)
        self.node_dim, self.edge_dim, self.keig, self.levels = node_dim, edge_dim, keig, levels
        self.sign_flip, self.generator, self.crop = sign_flip, generator, crop

    def __len__(self):
        return len(self.files)

    len = __len__

    def _fit(self, x, raw_dim):
        lead = raw_dim + (1 if self.levels > 1 else 0)              # + the cluster-id column of multi-level samples
        width = lead + self.keig - 1
        if x.shape[1] < width:
            x = torch.cat([x, torch.zeros(x.shape[0], width - x.shape[1], dtype=x.dtype)], dim=-1)
            if self.levels == 1:
                return x                                            # single-level get(): padded samples are not flipped (:430-431)
        else:
            x = x[:, :width]
        if not self.sign_flip:
            return x
        flips = -1 + 2 * torch.randint(0, 2, (self.keig - 1,), generator=self.generator)
        return x * torch.cat([torch.ones(lead), flips.to(x.dtype)])

    def get(self, idx):
        graph, _ = load_cached_graph(self.files[idx])
        if self.crop is not None:                                   # TSP_EigPE.get keeps the raw columns only (:693-694)
            graph.x_t, graph.x_s = graph.x_t[:, :self.crop[0]], graph.x_s[:, :self.crop[1]]
            return graph
        first = graph[0] if isinstance(graph, list) else graph
        first.x_t = self._fit(first.x_t, self.node_dim)
        first.x_s = self._fit(first.x_s, self.edge_dim)
        return graph

    __getitem__ = get

    def batch(self, indices, device=None):
        """`DataLoader` collation of the given samples; with `device` the tensors go there in one hop."""
        b = collate([self.get(i) for i in indices])
        if device is not None:
            for lv in (b if isinstance(b, list) else [b]):
                lv.to(device)
        return b


def _named(prefix, node_dim, edge_dim, levels=1, crop=None):
    def make(root, dataset=None, keig=8, num_pool=1, if_aug=False, **kw):
        ds = CachedGraphs(os.path.join(root, "processed") if os.path.isdir(os.path.join(root, "processed")) else root,
                          prefix, node_dim, edge_dim, keig, levels=levels, crop=crop, **kw)
        if dataset is not None:
            ds.files = ds.files[: len(dataset)]                     # the reference sizes the dataset by the raw one (:421-422)
        return ds
    return make


# the reference's dataset classes, `get` side only (process / download stay the reference's)
ZINC_HG_BM_par1_EigPE = _named("ZINC_BM_alleig_", 21, 3)                      # lib/Hodge_Dataset.py:408-440
ZINC_HG_BM_par1_MLGC = _named("ZINC_BM_MLGC_", 21, 3, levels=2)               # :480-515
Peptides_Func_EigPE = _named("Peptides_Func_alleig_", 9, 3)                   # :535-569
Peptides_Func_EigPE_MLGC = _named("Peptides_Func_alleig_MLGC_", 9, 3, levels=2)   # :608-650 (MLGC re-drawn per get there)
TSP_EigPE = _named("TSP_alleig_", 2, 1, crop=(2, 1))                          # :670-694
