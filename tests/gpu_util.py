"""Helpers shared by the -m gpu tests: NHWC/bf16 packing and thin ctypes call wrappers."""
import ctypes as C

import torch

from yolo_dbl_b200 import _lib as L


def dev():
    return torch.device("cuda:0")


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def to_nhwc(t_nchw: torch.Tensor, dtype=torch.bfloat16, pad_c: int = 0, c_off: int = 0):
    """NCHW fp32 (cpu) -> device NHWC buffer with `pad_c` extra channels; returns (buffer, dbl_tensor view)."""
    n, c, h, w = t_nchw.shape
    buf = torch.full((n, h, w, c + pad_c), 7.0, dtype=dtype, device=dev())  # poison the padding channels
    buf[..., c_off:c_off + c] = t_nchw.permute(0, 2, 3, 1).to(dev(), dtype)
    return buf, L.view(buf, c_off, c)


def empty_nhwc(n, h, w, c, dtype=torch.bfloat16, pad_c: int = 0, c_off: int = 0):
    buf = torch.full((n, h, w, c + pad_c), -3.0, dtype=dtype, device=dev())
    return buf, L.view(buf, c_off, c)


def from_nhwc(buf: torch.Tensor, c_off: int, c: int) -> torch.Tensor:
    return buf[..., c_off:c_off + c].permute(0, 3, 1, 2).float().cpu()


def stream():
    return L.current_stream()


def pack_w(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,k,k] fp32 -> device bf16 [Cout][k][k][Cin]"""
    return w.permute(0, 2, 3, 1).contiguous().to(dev(), torch.bfloat16)


def conv2d(x_view, w_dev, bias_dev, r_view, y_view, k, s, act, epi, impl, pad=None, dil=1):
    lib = L.require_device()
    rc = lib.dbl_conv2d(C.byref(x_view), w_dev.data_ptr(), None if bias_dev is None else bias_dev.data_ptr(),
                        None if r_view is None else C.byref(r_view), C.byref(y_view), k, s, k // 2 if pad is None else pad, dil, act,
                        epi, impl, stream())
    torch.cuda.synchronize()
    return rc


def assert_close_bf16(got: torch.Tensor, want: torch.Tensor, what: str, rel=2.0 ** -7, abs_=2e-2):
    """Output rounded once to bf16 (2^-9 relative) + fp32 accumulation-order noise."""
    err = (got - want).abs()
    tol = rel * want.abs() + abs_
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} outside tolerance, max err {float(err.max()):.4g} " \
                          f"(ref magnitude {float(want.abs().max()):.3g})"
