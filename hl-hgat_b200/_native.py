"""ctypes binding of libhlhgat.so (the C ABI declared in include/hlhgat.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception
is raised.  PyTorch is used only for device memory and streams -- every pointer handed to the
library is `tensor.data_ptr()` and every call is enqueued on `torch.cuda.current_stream()`.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhlhgat.so")

HL_TIE_POSITION, HL_TIE_COLUMN = 0, 1
HL_EPI_LAGUERRE_FIRST, HL_EPI_LAGUERRE_STEP, HL_EPI_CHEB_FIRST, HL_EPI_CHEB_STEP, HL_EPI_LINCOMB = range(5)
HL_LAGUERRE, HL_CHEB = 0, 1
HL_POST_NONE, HL_POST_CONST, HL_POST_RCP_ROW, HL_POST_MEAN = range(4)
HL_SIGMA_SIGMOID, HL_SIGMA_RELU = 0, 1
HL_MAX_SPMM_PROBLEMS = 4

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class SpmmProblem(C.Structure):
    _fields_ = [("rowptr", _vp), ("colidx", _vp), ("vals", _vp), ("nrows", _i32), ("nnz_hint", _i32),
                ("xg", _vp), ("ld_xg", _i64), ("p1", _vp), ("ld_p1", _i64), ("p2", _vp), ("ld_p2", _i64),
                ("p3", _vp), ("ld_p3", _i64), ("out", _vp), ("ld_out", _i64)]


class ConvSide(C.Structure):
    _fields_ = [("rowptr", _vp), ("colidx", _vp), ("vals", _vp), ("nrows", _i32), ("nnz_hint", _i32),
                ("x", _vp), ("ld_x", _i64), ("t", _vp), ("ld_t", _i64), ("t_stride", _i64), ("g0", _vp), ("ld_g0", _i64)]


class SplitDesc(C.Structure):
    _fields_ = [("src", _vp), ("hi", _vp), ("lo", _vp), ("ld_src", _i64), ("ld_out", _i64),
                ("rows", _i32), ("cols", _i32), ("transpose", _i32), ("reserved", _i32)]


class WgradReduceDesc(C.Structure):
    _fields_ = [("partial", _vp), ("cs_partial", _vp), ("dw", _vp), ("dw2", _vp), ("dbias", _vp),
                ("split_stride", _i64), ("ld_dw", _i64), ("ld_dw2", _i64),
                ("splits", _i32), ("fo", _i32), ("fi", _i32), ("fi_first", _i32), ("accumulate", _i32),
                ("accumulate_bias", _i32), ("block_start", _i32), ("reserved", _i32)]


class Hodge1Operator(C.Structure):
    _fields_ = [("inc_rowptr", _vp), ("inc_edge", _vp), ("tail", _vp), ("head", _vp), ("edge_scale", _vp),
                ("n_nodes", _i32), ("n_edges", _i32)]


_SIGNATURES = {
    "hl_version": (C.c_int, []),
    "hl_status_string": (C.c_char_p, [C.c_int]),
    "hl_last_cuda_error": (C.c_char_p, []),
    "hl_device_sm_count": (C.c_int, []),
    "hl_launch_count": (C.c_ulonglong, []),
    "hl_csr_from_coo_workspace": (_sz, [_i64, _i64]),
    "hl_csr_from_coo": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, C.c_int, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hl_set_spmm_mode": (None, [C.c_int]),
    "hl_poly_spmm": (C.c_int, [C.POINTER(SpmmProblem), C.c_int, _i32, C.c_int, C.POINTER(_f32), _vp]),
    "hl_poly_basis_fwd": (C.c_int, [C.c_int, C.c_int, C.POINTER(ConvSide), C.c_int, _i32, _vp]),
    "hl_poly_basis_bwd": (C.c_int, [C.c_int, C.c_int, C.POINTER(ConvSide), C.c_int, _i32, _vp]),
    "hl_segment_reduce": (C.c_int, [_vp, _vp, _i32, _vp, _i64, _vp, _vp, _i64, _i32, C.c_int, _vp, _f32, _vp]),
    "hl_endpoint_gather": (C.c_int, [_vp, _vp, _i32, _vp, _i64, _vp, _vp, _i64, _i32, _f32, _vp]),
    "hl_boundary_absdiff_fwd": (C.c_int, [_vp, _vp, _i32, _vp, _i64, _vp, _i64, _i32, _f32, _vp]),
    "hl_boundary_absdiff_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _f32, _vp]),
    "hl_owner_gather": (C.c_int, [_vp, _i32, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp]),
    "hl_att_gate_fwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _f32, C.c_int, _vp, _vp]),
    "hl_att_gate_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, C.c_int, _vp, _vp, _vp, _vp]),
    "hl_build_edges_workspace": (_sz, [_i64]),
    "hl_build_edges": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hl_lambda_max_workspace": (_sz, [_i32, _i32, _i32]),
    "hl_lambda_max": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "hl_laplacian_rowptr_workspace": (_sz, [_i32, _i32]),
    "hl_laplacian_rowptr": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hl_laplacian_fill": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hl_tf32_split": (C.c_int, [_vp, _i64, _i32, _i32, C.c_int, _vp, _vp, _i64, _vp]),
    "hl_tf32_split_batch": (C.c_int, [_vp, _i32, _i64, _vp]),
    "hl_gemm_tf32x3": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _i64, C.c_int, _vp]),
    "hl_gemm2_tf32x3": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _i64, C.c_int, _vp]),
    "hl_gemm_bn_part_floats": (_sz, [_i32, _i32]),
    "hl_gemm2_bn_tf32x3": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _i64, C.c_int, _vp, _vp, _vp]),
    "hl_wgrad_tf32x3_workspace": (_sz, [_i32, _i32, _i32]),
    "hl_wgrad_tf32x3": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _i64, C.c_int, _vp, _sz, _vp]),
    "hl_wgrad_bias_tf32x3": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _i64, C.c_int, _vp, C.c_int, _vp, _sz, _vp]),
    "hl_wgrad2_tf32x3_workspace": (_sz, [_i32, _i32, _i32]),
    "hl_wgrad2_bias_tf32x3": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _vp, _i64, C.c_int, _vp, C.c_int,
                                        _vp, _sz, _vp]),
    "hl_wgrad_deferred_tf32x3": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _vp, _i64, C.c_int, _vp, C.c_int,
                                           _vp, _sz, C.POINTER(WgradReduceDesc), _vp]),
    "hl_wgrad_reduce_batch": (C.c_int, [C.POINTER(WgradReduceDesc), _i32, _vp]),
    "hl_wgrad_workspace": (_sz, [_i32, _i32, _i32]),
    "hl_wgrad": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _i64, C.c_int, _vp, _sz, _vp]),
    "hl_colsum_workspace": (_sz, [_i32, _i32]),
    "hl_colsum": (C.c_int, [_vp, _i64, _i32, _i32, _vp, C.c_int, _vp, _sz, _vp]),
    "hl_poly_basis_hodge1_fwd": (C.c_int, [C.c_int, C.c_int, C.POINTER(Hodge1Operator), _vp, _i64, _vp, _i64, _i64, _vp, _i32, _vp]),
    "hl_poly_basis_hodge1_bwd": (C.c_int, [C.c_int, C.c_int, C.POINTER(Hodge1Operator), _vp, _i64, _vp, _i64, _i64, _vp, _i32, _vp]),
    "hl_adam_flat": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "hl_eig_pe_workspace": (_sz, [_i64]),
    "hl_eig_pe": (C.c_int, [_vp, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _vp, _i32, _vp, _sz, _vp]),
    "hl_greedy_matching": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hl_bn_workspace": (_sz, [_i32, _i32]),
    "hl_bn_act_fwd": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _f32, _f32, _vp, _i64, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _sz, _vp]),
    "hl_bn_act_fwd_tiles": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _f32, _f32, _vp, _i64, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp]),
    "hl_bn_act_bwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _f32, _f32,
                                _vp, _i64, _vp, _vp, C.c_int, _vp, _vp, _sz, _vp]),
    "hl_bn_stats": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "hl_bn_apply": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _f32, _f32, _vp, _i64, _vp, _vp]),
    "hl_bn_bwd_sums": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _f32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "hl_bn_bwd_apply": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _i64, _vp, _vp]),
}

_lib = None


class HlError(RuntimeError):
    pass


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HlError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the ABI and the binding drifted apart
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(status, what):
    if status != 0:
        L = lib()
        msg = L.hl_status_string(status).decode()
        if status == -3:
            msg += ": " + L.hl_last_cuda_error().decode()
        raise HlError(f"{what} failed: {msg}")


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def require_cuda_f32(*tensors):
    """All tensors CUDA fp32 and on the CURRENT device: every call is enqueued on `torch.cuda.current_stream()` and the
    kernels' one-time configuration is looked up for the current device (one process per GPU; a model on `cuda:1`
    needs `torch.cuda.set_device(1)` or a `with torch.cuda.device(1):` around its calls)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise HlError("hlhgat_b200 ops run on CUDA tensors only (no CPU fallback)")
        if t.dtype != torch.float32:
            raise HlError(f"expected float32, got {t.dtype}")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise HlError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: run under "
                          f"torch.cuda.device({t.device.index}) (kernels are enqueued on the current device's stream)")


def row_major(t):
    """2-D tensor with unit column stride; returns (tensor, leading dimension)."""
    if t.dim() != 2:
        raise HlError("expected a 2-D tensor")
    if t.stride(1) != 1 and t.shape[1] > 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0)))
