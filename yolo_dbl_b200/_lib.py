"""ctypes binding of libdbl_sm100.so (the C ABI declared in include/dbl_sm100.h).

There is no fallback of any kind: if the shared library is missing or the device is not sm_100, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdbl_sm100.so")

DBL_BF16, DBL_F32 = 0, 1
ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2
EPI_NONE, EPI_ADD, EPI_MUL = 0, 1, 2
IMPL_AUTO, IMPL_DIRECT, IMPL_TCGEN05, IMPL_TCGEN05_TMA, IMPL_TCGEN05_MULTI = 0, 1, 2, 3, 4


class DblError(RuntimeError):
    pass


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("pitch", C.c_int32), ("dtype", C.c_int32)]


class HypergraphWeights(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("prototype_base", "context_w", "context_b", "pre_head_w", "pre_head_b",
                                          "edge_w", "edge_b", "node_w", "node_b")] + [("num_edges", C.c_int32),
                                                                                        ("num_heads", C.c_int32)]


class HeadLevels(C.Structure):
    _fields_ = [("data", C.c_void_p * 3), ("h", C.c_int32 * 3), ("w", C.c_int32 * 3), ("stride", C.c_float * 3),
                ("num_levels", C.c_int32), ("batch", C.c_int32), ("nc", C.c_int32), ("pitch", C.c_int32)]


class HeadFusedLevel(C.Structure):
    _fields_ = [("f2", C.c_void_p), ("f3", C.c_void_p), ("w2_packed", C.c_void_p), ("w3_packed", C.c_void_p), ("b2", C.c_void_p),
                ("b3", C.c_void_p), ("logits", C.c_void_p), ("pitch2", C.c_int32), ("pitch3", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("stride", C.c_float)]


class HeadFused(C.Structure):
    _fields_ = [("level", HeadFusedLevel * 3), ("num_levels", C.c_int32), ("batch", C.c_int32), ("nc", C.c_int32),
                ("c2", C.c_int32), ("c3", C.c_int32), ("logit_pitch", C.c_int32)]


class LbImage(C.Structure):
    _fields_ = [("src", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("pitch", C.c_int32), ("new_h", C.c_int32), ("new_w", C.c_int32),
                ("top", C.c_int32), ("left", C.c_int32), ("reserved", C.c_int32)]


class BoxScale(C.Structure):
    _fields_ = [("pad_x", C.c_float), ("pad_y", C.c_float), ("gain", C.c_float), ("w0", C.c_float), ("h0", C.c_float)]


class Candidates(C.Structure):
    _fields_ = [("box", C.c_void_p), ("score", C.c_void_p), ("cls", C.c_void_p), ("pos", C.c_void_p),
                ("count", C.c_void_p), ("cap", C.c_int32), ("batch", C.c_int32)]


_P = C.c_void_p
_TP = C.POINTER(Tensor)
_SIGNATURES = {
    "dbl_last_error": (C.c_char_p, []),
    "dbl_version": (C.c_int, []),
    "dbl_device_ok": (C.c_int, []),
    "dbl_set_pdl": (C.c_int, [C.c_int]),
    "dbl_launch_count": (C.c_int64, []),
    "dbl_conv2d": (C.c_int, [_TP, _P, _P, _TP, _TP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "dbl_conv2d_which": (C.c_int, [_TP, _TP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dbl_stem_conv": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _TP, C.c_int, _P]),
    "dbl_stem_conv_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _TP, C.c_int, _P]),
    "dbl_mma_packed_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dbl_mma_pack_weights": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "dbl_bottleneck3x3": (C.c_int, [_TP, _P, _P, _P, _P, C.c_int, _TP, _P]),
    "dbl_stem_tc_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dbl_stem_tc_packed_bytes": (C.c_size_t, []),
    "dbl_stem_tc_pack": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P]),
    "dbl_stem_pair_tc": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _TP, _P]),
    "dbl_bottleneck_tc_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "dbl_bottleneck_tc_packed_bytes": (C.c_size_t, [C.c_int]),
    "dbl_bottleneck_tc_pack": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P]),
    "dbl_bottleneck3x3_tc": (C.c_int, [_TP, _P, C.c_int, _TP, _P]),
    "dbl_conv3x3_mma": (C.c_int, [_TP, _P, _P, _TP, _TP, C.c_int, C.c_int, C.c_int, _P]),
    "dbl_stem_pair": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, _TP, _P]),
    "dbl_dsconv_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dbl_dsconv": (C.c_int, [_TP, _P, _P, C.c_int, _P, _P, _TP, _TP, C.c_int, C.c_int, C.c_int, _P]),
    "dbl_dsconv_tc_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dbl_dsconv_tc": (C.c_int, [_TP, _P, _P, C.c_int, _P, _P, _TP, _TP, C.c_int, C.c_int, _P]),
    "dbl_dsconv_tc_debug_trace": (C.c_int, [_P]),
    "dbl_area_attention": (C.c_int, [_TP, _TP, _TP, C.c_int, C.c_int, _P]),
    "dbl_letterbox_u8": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P]),
    "dbl_dwconv2d": (C.c_int, [_TP, _P, _P, _TP, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "dbl_avgpool2": (C.c_int, [_TP, _TP, _P]),
    "dbl_upsample2_nearest": (C.c_int, [_TP, _TP, _P]),
    "dbl_gated_add": (C.c_int, [_TP, _TP, C.c_float, _TP, _TP, _P]),
    "dbl_copy": (C.c_int, [_TP, _TP, _P]),
    "dbl_channel_scale_add": (C.c_int, [_TP, _TP, _P, _TP, _P]),
    "dbl_dysample": (C.c_int, [_TP, _TP, _TP, C.c_int, _P]),
    "dbl_lsk_squeeze": (C.c_int, [_TP, _TP, _P]),
    "dbl_lsk_gate": (C.c_int, [_TP, _TP, _P, _P, _TP, _P]),
    "dbl_hypergraph_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dbl_hypergraph": (C.c_int, [_TP, C.POINTER(HypergraphWeights), _TP, _P, C.c_size_t, _P]),
    "dbl_detect_decode": (C.c_int, [C.POINTER(HeadLevels), _P, _P]),
    "dbl_nms_filter": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_int, _P,
                                 C.c_int, C.POINTER(Candidates), _P]),
    "dbl_decode_filter": (C.c_int, [C.POINTER(HeadLevels), C.c_float, C.c_int, _P, C.c_int, C.POINTER(Candidates), _P]),
    "dbl_head_decode_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "dbl_head_decode": (C.c_int, [C.POINTER(HeadFused), _P, C.c_float, C.c_int, _P, C.c_int, C.POINTER(Candidates), _P]),
    "dbl_nms_sort_capacity": (C.c_size_t, [C.c_int]),
    "dbl_nms": (C.c_int, [C.POINTER(Candidates), C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "dbl_nms_scaled": (C.c_int, [C.POINTER(Candidates), C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "dbl_nms_append_labels": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, C.c_int, C.POINTER(Candidates), _P]),
    "dbl_match_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dbl_match_predictions": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P, C.c_size_t, _P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (no CUDA call is made).  Raises DblError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DblError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           f"(or `make -C yolo_dbl_b200/csrc`).  There is no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if os.environ.get("DBL_PDL", "1") == "0":  # A/B switch for programmatic dependent launch
            lib.dbl_set_pdl(0)
        _lib = lib
    return _lib


def require_device() -> C.CDLL:
    lib = load()
    if not torch.cuda.is_available():
        raise DblError("libdbl_sm100 needs an sm_100 (B200) GPU: torch.cuda.is_available() is False and there is no fallback")
    if not lib.dbl_device_ok():
        raise DblError(f"unsupported device: {lib.dbl_last_error().decode()}")
    return lib


def check(code: int, what: str = ""):
    if code != 0:
        raise DblError(f"{what} failed with code {code}: {load().dbl_last_error().decode()}")


def current_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_DT = {torch.bfloat16: DBL_BF16, torch.float32: DBL_F32}


def view(t: torch.Tensor, c_off: int = 0, c: Optional[int] = None) -> Tensor:
    """dbl_tensor describing channels [c_off, c_off+c) of a dense NHWC buffer t[N,H,W,Cbuf]."""
    assert t.dim() == 4 and t.is_contiguous(), "expected a contiguous NHWC buffer"
    n, h, w, cb = t.shape
    c = cb - c_off if c is None else c
    assert 0 <= c_off and c_off + c <= cb
    return Tensor(t.data_ptr() + c_off * t.element_size(), n, h, w, c, cb, _DT[t.dtype])


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()
