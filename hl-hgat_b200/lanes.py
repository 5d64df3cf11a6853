"""Node lane / edge lane: the two halves of an HL-HGAT layer on two CUDA streams.

Inside a layer the node chain (`conv_t -> BN -> act`, the node MLP of `NodeEdgeInt`) and the edge chain
(`conv_s -> BN -> act`, the edge MLP) are independent (lib/Hodge_Cheb_Conv.py:142-154, :306-309); they meet only
at the simplex transfers (`|B1| x_s` needs the edge features, `|B1|^T x_t` the node features, :294-295), at the
pooling block and at the readout.  At ZINC batch sizes one kernel does not fill 148 SMs (GEMM grids of 1.3 waves,
BatchNorm finalize kernels of a few CTAs), so the model classes issue everything that produces edge features on
a second stream: captured into the whole-step CUDA graph this becomes two parallel branches, and the tail of one
lane's kernel overlaps the other lane's work.  Autograd replays every backward node on the stream its forward ran
on and orders producer / consumer streams itself, so the backward pass is two-lane as well.

Arithmetic is untouched -- the same kernels run in the same per-lane order -- so the forward pass is bit-identical
with the lanes on or off.  In the backward pass autograd adds the contributions that reach one tensor from two
streams in a different association than on one stream, so gradients agree to fp32 rounding (~1e-7 relative;
tools/lanes_debug.py shows the same differences under CUDA_LAUNCH_BLOCKING=1, i.e. it is summation order, not a
race) and two-lane runs are bit-identical to each other (tests/test_gpu_lanes.py).

Memory safety across streams: a tensor allocated on one lane and read on the other is `record_stream`-ed at the
crossing, so the caching allocator never hands its block to the allocating lane while the other lane may still
read it (under graph capture such blocks are simply not reused before the capture ends).
"""
import contextlib

import torch

_STATE = {"enabled": False, "active": None, "dirty": False}
_SIDE = {}          # device index -> the edge-lane stream (created once; a captured graph keeps using it)
_AUX = {}           # (device index, lane stream handle) -> that lane's weight-gradient stream
_AUX_USED = []      # aux streams with work issued since the last join()


def enable_lanes(flag=True):
    """Opt in: model classes of lib/Hodge_ST_Model.py run the edge chain on a second stream."""
    _STATE["enabled"] = bool(flag)


def lanes_enabled():
    return _STATE["enabled"]


def active():
    """The `Lanes` of the forward pass being issued, or None (single stream)."""
    return _STATE["active"]


class Lanes:
    def __init__(self, device):
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.node = torch.cuda.current_stream(idx)
        side = _SIDE.get(idx)
        if side is None:
            side = _SIDE[idx] = torch.cuda.Stream(device=idx)
        self.edge = side

    def edge_ctx(self):
        """`with lanes.edge_ctx():` -- issue on the edge lane."""
        return torch.cuda.stream(self.edge)

    @staticmethod
    def _mark(stream, tensors):
        for t in tensors:
            if t is not None and t.is_cuda:
                t.record_stream(stream)

    def to_edge(self, *tensors):
        """The edge lane waits for everything issued so far on the node lane; `tensors` (allocated on the node
        lane) are about to be read on the edge lane."""
        self.edge.wait_stream(self.node)
        self._mark(self.edge, tensors)

    def to_node(self, *tensors):
        self.node.wait_stream(self.edge)
        self._mark(self.node, tensors)

    def exchange(self, node_tensors=(), edge_tensors=()):
        """Both lanes wait for what the other has issued up to this point (and for nothing issued later):
        `node_tensors` are about to be read on the edge lane, `edge_tensors` on the node lane."""
        ev_n, ev_e = torch.cuda.Event(), torch.cuda.Event()
        ev_n.record(self.node)
        ev_e.record(self.edge)
        self.node.wait_event(ev_e)
        self.edge.wait_event(ev_n)
        self._mark(self.edge, node_tensors)
        self._mark(self.node, edge_tensors)


@contextlib.contextmanager
def open_lanes(device):
    """Context of one forward pass.  Yields a `Lanes` (edge lane already waiting for the node lane) when lanes are
    enabled and none is open yet, else None.  On exit the node lane waits for the edge lane."""
    if not _STATE["enabled"] or _STATE["active"] is not None or torch.device(device).type != "cuda":
        yield None
        return
    ln = Lanes(device)
    ln.edge.wait_stream(ln.node)
    _STATE["active"] = ln
    _STATE["dirty"] = True
    try:
        yield ln
    finally:
        _STATE["active"] = None
        ln.node.wait_stream(ln.edge)


@contextlib.contextmanager
def weight_grad_lane(*tensors):
    """Backward pass, fused-accumulation mode only (functional.accumulate_into_grads): weight / bias gradients are
    added straight into the flat bucket and nothing downstream in the autograd graph reads them, so they leave the
    critical chain (data gradient -> BatchNorm backward -> ...) and are issued on an auxiliary stream of the current
    lane -- one more parallel branch of the whole-step graph, ordered after everything issued so far on the lane
    and joined by `join()`.  `tensors` (the operands the side stream reads) are `record_stream`-ed."""
    if not _STATE["enabled"]:
        yield
        return
    cur = torch.cuda.current_stream()
    key = (cur.device.index, cur.cuda_stream)
    aux = _AUX.get(key)
    if aux is None:
        aux = _AUX[key] = torch.cuda.Stream(device=cur.device)
    aux.wait_stream(cur)
    Lanes._mark(aux, tensors)
    if aux not in _AUX_USED:
        _AUX_USED.append(aux)
    _STATE["dirty"] = True
    with torch.cuda.stream(aux):
        yield


def join(device=None):
    """After `loss.backward()`: the current stream waits for the edge lane (weight gradients accumulated straight
    into the flat bucket bypass autograd's own end-of-backward stream sync)."""
    if not _STATE["dirty"]:
        return
    _STATE["dirty"] = False
    idx = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
    cur = torch.cuda.current_stream(idx)
    side = _SIDE.get(idx)
    if side is not None:
        cur.wait_stream(side)
    for aux in _AUX_USED:
        cur.wait_stream(aux)
    _AUX_USED.clear()
