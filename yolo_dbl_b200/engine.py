"""Lowering of the YOLO-DBL layer table to a static list of libdbl_sm100 launches (the B200 executor).

One ``DBLEngine`` is built per (weights, batch, height, width).  Building it
  * folds every BatchNorm into the preceding convolution (including DSConv's, which the reference leaves unfolded,
    ``nn/tasks.py:217``) and repacks weights to the kernels' layouts (bf16 ``[Cout][kh][kw][Cin]`` for the implicit
    GEMM, f32 ``[kh*kw][C]`` for depthwise),
  * allocates every activation once (NHWC bf16; concat / chunk are channel slices of shared buffers, so the
    reference's ``torch.cat`` copies disappear),
  * records the launches as plain ctypes calls, which are then captured into a CUDA graph.
Replays are sync-free: fixed-capacity outputs + device-side counts.

Reference control flow being replaced: ``BaseModel._predict_once`` (``nn/tasks.py:145-172``) and the module
``forward`` methods cited next to each ``_lower_*`` function.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .graph import GraphSpec, LayerSpec, detect_channels

BN_EPS = 1e-3  # utils/torch_utils.py:410-420


class Act:
    """Channel slice [off, off+c) of a dense NHWC device buffer."""

    __slots__ = ("buf", "off", "c", "_desc")

    def __init__(self, buf: torch.Tensor, off: int = 0, c: Optional[int] = None):
        self.buf, self.off = buf, off
        self.c = buf.shape[3] - off if c is None else c
        self._desc = None

    n = property(lambda s: s.buf.shape[0])
    h = property(lambda s: s.buf.shape[1])
    w = property(lambda s: s.buf.shape[2])

    @property
    def desc(self) -> L.Tensor:
        if self._desc is None:
            self._desc = L.view(self.buf, self.off, self.c)
        return self._desc

    def slice(self, off: int, c: int) -> "Act":
        assert off + c <= self.c
        return Act(self.buf, self.off + off, c)

    def nchw(self) -> torch.Tensor:
        """fp32 NCHW copy (tests / API compatibility)."""
        return self.buf[..., self.off:self.off + self.c].permute(0, 3, 1, 2).float().contiguous().cpu()


def _fold(w: torch.Tensor, sd, bn: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """fuse_conv_and_bn (utils/torch_utils.py:238-265) for a bias-free conv."""
    scale = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + BN_EPS)
    return w.float() * scale.view(-1, 1, 1, 1), sd[bn + ".bias"].float() - sd[bn + ".running_mean"].float() * scale


class DBLEngine:
    def __init__(self, spec: GraphSpec, state_dict: Dict[str, torch.Tensor], batch: int, height: int, width: int,
                 device: torch.device | str = "cuda", conv_impl: int = L.IMPL_AUTO, use_graph: bool = True,
                 keep_layers: bool = False, fuse_thin: Optional[bool] = None, lanes: int = 8):
        self.lib = L.require_device()
        self.spec, self.B, self.H, self.W = spec, int(batch), int(height), int(width)
        if self.H % 32 or self.W % 32:
            raise ValueError("input height/width must be multiples of 32 (reference data/loaders.py:516-584)")
        self.dev = torch.device(device)
        self.conv_impl = conv_impl
        # warp-MMA kernels for the thin high-resolution layers (fused stem pair, fused Bottleneck, thin stride-2 convs);
        # off when a specific conv implementation is forced or every top-level layer output must exist (parity tests)
        self.fuse_thin = (conv_impl == L.IMPL_AUTO) if fuse_thin is None else bool(fuse_thin)
        self.fuse_stem = self.fuse_thin and not keep_layers
        self.fuse_head = self.fuse_thin and not keep_layers  # last head convs + decode + filter in one kernel (dbl_head_decode)
        # dw + pw of a DSConv / Detect class-branch pair in one kernel (dbl_dsconv).  "auto" (default): only where it was
        # measured faster than the two-kernel form -- stride 1 and a small problem (H*W*Cin per image <= 128 Ki: the
        # DSBottleneck halves at 40^2 / 20^2 and the small Detect class-branch pairs), where both launches are mostly fixed
        # cost.  Larger maps stay unfused: the fused CTA serialises fill -> depthwise -> pointwise at low occupancy
        # (Detect P3 64->80 @80^2: 108 us fused vs 89 us as two launches).  "1" = wherever supported, "0" = never.
        self.fuse_ds = os.environ.get("DBL_FUSE_DS", "auto")
        self.ds_impl = os.environ.get("DBL_DS_IMPL", "auto")  # "mma": keep the warp-MMA kernel (A/B aid); "auto": tcgen05 kernel where instantiated
        # layers 0+1 on the tcgen05 pixel-group kernel (csrc/stem_tc.cu): measured SLOWER than the warp-MMA stem (280 vs 253 us on f32 input, 237 vs
        # ~250 on u8: bound by shared-memory bandwidth), so it is opt-in
        self.stem_tc = os.environ.get("DBL_STEM_TC", "0") == "1"
        self.bn_tc = os.environ.get("DBL_BN_TC", "1") != "0"  # thin Bottlenecks (C = 16, 32) on the tcgen05 pixel-group kernel; "0": round-1 kernels (A/B aid)
        self.fuse_bn_max_c = int(os.environ.get("DBL_FUSE_BN_MAXC", "16"))  # widest Bottleneck that uses the fused warp-MMA kernel
        self.use_graph = use_graph
        # independent branches of the graph (Detect's six conv chains, HyperACE's three branches, the neck's side paths)
        # run on `lanes` streams: inside a captured CUDA graph they become parallel branches, so small kernels overlap
        self.n_lanes = max(1, int(os.environ.get("DBL_LANES", lanes))) if use_graph else 1
        self.sd = {k: v.detach().to("cpu") for k, v in state_dict.items()}
        self._keep: List[object] = []        # weights, ctypes structs, scratch: everything launches point at
        self.ops: List[Tuple[Callable, tuple, str]] = []
        self.op_rw: List[Tuple[list, list]] = []
        self.op_meta: List[dict] = []
        self.kernels_per_step: Dict[tuple, int] = {}
        self.layer_out: List[Optional[Act]] = []
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self._post_bufs: Dict[tuple, dict] = {}
        with torch.cuda.device(self.dev):
            self.x_in = torch.zeros(self.B, 3, self.H, self.W, dtype=torch.float32, device=self.dev)
            self.x_in_u8 = torch.zeros(self.B, 3, self.H, self.W, dtype=torch.uint8, device=self.dev)  # image sources (0..255)
            self._u8 = False
            self._build()
            self._plan_lanes()
        if not keep_layers:
            self.sd = None

    # ------------------------------------------------------------------------------------------------ helpers
    def _alloc(self, h: int, w: int, c: int, dtype=torch.bfloat16) -> Act:
        buf = torch.zeros(self.B, h, w, c, dtype=dtype, device=self.dev)
        self._keep.append(buf)  # launches hold raw pointers: the engine owns every buffer for its whole life
        return Act(buf)

    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        t = t.to(dtype).contiguous().to(self.dev)
        self._keep.append(t)
        return t

    def _emit(self, fn, args: tuple, what: str, kern: str = "", flops: float = 0.0, nbytes: float = 0.0):
        """Record one launch.  kern / flops / nbytes = kernel family and ALGORITHMIC work: every
        declared input and output tensor counted once, weights once; halo re-reads and scratch are not counted."""
        self._keep.append(args)
        self.ops.append((fn, args, what))
        # memory regions of the tensor arguments, for the dependency analysis of _plan_lanes: by convention of the C ABI
        # the LAST dbl_tensor argument is the output (dbl_gated_add: both tensors after the gate), the others are inputs
        views = [(i, a._obj) for i, a in enumerate(args) if isinstance(getattr(a, "_obj", None), L.Tensor)]
        n_out = 1
        if fn is self.lib.dbl_gated_add:
            n_out = len([1 for i, _ in views if i >= 3])
        elif fn is self.lib.dbl_area_attention:
            n_out = 2  # (qkv, o, v_out): o and v_out are written
        regs = [self._region(t) for _, t in views]
        self.op_rw.append((regs[:len(regs) - n_out], regs[len(regs) - n_out:]))
        if not nbytes:  # bandwidth-bound glue: every tensor view in the argument list, once
            for a in args:
                t = getattr(a, "_obj", None)
                if isinstance(t, L.Tensor):
                    nbytes += t.n * t.h * t.w * t.c * (2 if t.dtype == L.DBL_BF16 else 4)
            if fn is self.lib.dbl_stem_conv:
                nbytes += self.x_in.numel() * 4
        self.op_meta.append(dict(what=what, kern=kern or fn.__name__, flops=float(flops), bytes=float(nbytes)))

    def _pack_w(self, w: torch.Tensor) -> torch.Tensor:
        """[Cout,Cin,kh,kw] f32 -> device bf16 [Cout][kh][kw][Cin]"""
        return self._dev(w.permute(0, 2, 3, 1), torch.bfloat16)

    def _conv(self, x: Act, w: torch.Tensor, b: Optional[torch.Tensor], y: Act, k: int, s: int, act: int,
              epi: int = L.EPI_NONE, r: Optional[Act] = None, what: str = "conv"):
        wd = self._pack_w(w)
        bd = self._dev(b, torch.float32) if b is not None else None
        args = (C.byref(x.desc), L.ptr(wd), L.ptr(bd), C.byref(r.desc) if r is not None else None, C.byref(y.desc), k, s,
                k // 2, 1, act, epi, self.conv_impl)
        npix = self.B * y.h * y.w
        ysz = 2 if y.buf.dtype == torch.bfloat16 else 4
        nbytes = 2 * self.B * x.h * x.w * x.c + ysz * npix * y.c + 2 * w.numel() + 4 * y.c + (2 * npix * y.c if r is not None else 0)
        which = self.lib.dbl_conv2d_which(C.byref(x.desc), C.byref(y.desc), k, s, k // 2, 1, self.conv_impl)
        kern = {0: "conv_direct", 1: "conv_tc_1x1" if k == 1 else "conv_tc_tma", 3: "conv_tc3", 4: "conv_tc4", 5: "conv_s2_tc", 6: "conv_tc5"}.get(which, "conv_direct")
        self._emit(self.lib.dbl_conv2d, args, what, kern, 2.0 * npix * y.c * x.c * k * k, nbytes)

    def _pack_mma(self, w: torch.Tensor) -> torch.Tensor:
        """[Cout,Cin,3,3] f32 -> device buffer in MMA B-fragment order (dbl_mma_pack_weights)"""
        cout, cin, kh, kw = w.shape
        wd = self._pack_w(w)
        nbytes = self.lib.dbl_mma_packed_bytes(cout, cin, kh * kw)
        if not nbytes:
            raise L.DblError(f"no warp-MMA packing for cout={cout} cin={cin}")
        out = torch.zeros(nbytes // 4, dtype=torch.int32, device=self.dev)
        L.check(self.lib.dbl_mma_pack_weights(L.ptr(wd), cout, cin, kh * kw, L.ptr(out), L.current_stream()), "dbl_mma_pack_weights")
        torch.cuda.synchronize(self.dev)
        self._keep.append(out)
        return out

    THIN_S2 = {(8, 16), (16, 32), (32, 64)}

    def _conv_bn_thin(self, x: Act, p: str, y: Act, s: int) -> bool:
        """3x3 Conv module on the warp-MMA kernel when an instantiation exists; False -> caller uses dbl_conv2d"""
        w, b = _fold(self.sd[p + ".conv.weight"], self.sd, p + ".bn")
        if not (self.fuse_thin and w.shape[-1] == 3 and s == 2 and (x.c, y.c) in self.THIN_S2 and x.buf.data_ptr() != y.buf.data_ptr()):
            return False
        if self.lib.dbl_conv2d_which(C.byref(x.desc), C.byref(y.desc), 3, s, 1, 1, self.conv_impl) == 5:
            return False  # dbl_conv2d runs it on the tcgen05 pixel-pair kernel (conv_s2_tc.cu)
        wp, bd = self._pack_mma(w), self._dev(b, torch.float32)
        npix = self.B * y.h * y.w
        self._emit(self.lib.dbl_conv3x3_mma, (C.byref(x.desc), L.ptr(wp), L.ptr(bd), None, C.byref(y.desc), s, L.ACT_SILU, L.EPI_NONE), p,
                   "conv_mma", 2.0 * npix * y.c * x.c * 9, 2 * self.B * x.h * x.w * x.c + 2 * npix * y.c + 2 * w.numel() + 4 * y.c)
        return True

    def _conv_bn(self, x: Act, p: str, y: Act, s: int = 1, act: int = L.ACT_SILU, epi=L.EPI_NONE, r=None, g: int = 1):
        """Conv module: conv(bias=False)+BN+SiLU (nn/modules/conv.py:39-63).  Grouped convolutions (g > 1: the strided stem of the stock
        yolov13.yaml) run as g dense convolutions on channel slices of x and y -- no copy, no extra kernel."""
        w, b = _fold(self.sd[p + ".conv.weight"], self.sd, p + ".bn")
        if g == 1:
            return self._conv(x, w, b, y, w.shape[-1], s, act, epi, r, what=p)
        ci, co = x.c // g, y.c // g
        for gi in range(g):
            self._conv(x.slice(gi * ci, ci), w[gi * co:(gi + 1) * co], b[gi * co:(gi + 1) * co], y.slice(gi * co, co), w.shape[-1], s, act, epi,
                       r.slice(gi * co, co) if r is not None else None, what=f"{p}.g{gi}")

    def _conv_bn_pair(self, x: Act, p1: str, p2: str, y: Act):
        """two 1x1 Conv modules reading the same input as ONE GEMM writing [p1 | p2] (C3-style cv1 / cv2)"""
        w1, b1 = _fold(self.sd[p1 + ".conv.weight"], self.sd, p1 + ".bn")
        w2, b2 = _fold(self.sd[p2 + ".conv.weight"], self.sd, p2 + ".bn")
        self._conv(x, torch.cat([w1, w2], 0), torch.cat([b1, b2], 0), y, 1, 1, L.ACT_SILU, what=f"{p1}|cv2")

    def _conv_bias(self, x: Act, p: str, y: Act, epi=L.EPI_NONE, r=None):
        """plain nn.Conv2d 1x1 with bias, no activation"""
        self._conv(x, self.sd[p + ".weight"].float(), self.sd[p + ".bias"].float(), y, 1, 1, L.ACT_NONE, epi, r, what=p)

    def _dw(self, x: Act, w: torch.Tensor, b: Optional[torch.Tensor], y: Act, k: int, s: int, d: int, act: int, what: str):
        wd = self._dev(w.reshape(w.shape[0], k * k).t(), torch.float32)  # [C,1,k,k] -> [k*k][C]
        bd = self._dev(b, torch.float32) if b is not None else None
        npix = self.B * y.h * y.w
        self._emit(self.lib.dbl_dwconv2d, (C.byref(x.desc), L.ptr(wd), L.ptr(bd), C.byref(y.desc), k, s, d, act), what, "dwconv",
                   2.0 * npix * y.c * k * k, 2 * self.B * x.h * x.w * x.c + 2 * npix * y.c + 4 * k * k * y.c)

    def _dsconv_fused(self, x: Act, wdw: torch.Tensor, bdw: Optional[torch.Tensor], wpw: torch.Tensor, bpw: torch.Tensor, y: Act,
                      s: int, epi, r, what: str) -> bool:
        """dw k x k (+bias+SiLU when bdw is given) -> pw 1x1 + bias + SiLU (+ residual) in one kernel (csrc/dsconv_mma.cu);
        False when no instantiation exists or fusion is off -> the caller emits the two-kernel form"""
        k, cin, cout = wdw.shape[-1], wdw.shape[0], wpw.shape[0]
        # Blackwell data path first (csrc/dsconv_tc.cu: TMA + tcgen05 + TMEM, persistent, phases of consecutive tiles overlap): faster
        # than both the warp-MMA kernel and the two-launch form on every shape measured (tools/ds_bench.py, 1.2x .. 2.3x)
        if (self.fuse_thin and self.ds_impl != "mma" and self.fuse_ds != "0" and x.buf.data_ptr() != y.buf.data_ptr()
                and not (bdw is not None and r is not None) and self.lib.dbl_dsconv_tc_supported(k, s, cin, cout)):
            wd = self._dev(wdw.reshape(cin, k * k).t(), torch.float32)
            bd = self._dev(bdw, torch.float32) if bdw is not None else None
            wp, bp = self._pack_w(wpw), self._dev(bpw, torch.float32)
            npix = self.B * y.h * y.w
            args = (C.byref(x.desc), L.ptr(wd), L.ptr(bd), 1 if bdw is not None else 0, L.ptr(wp), L.ptr(bp),
                    C.byref(r.desc) if r is not None else None, C.byref(y.desc), k, epi if r is not None else L.EPI_NONE)
            nbytes = 2 * self.B * x.h * x.w * cin + 2 * npix * cout * (2 if r is not None else 1) + 4 * k * k * cin + 2 * cin * cout
            self._emit(self.lib.dbl_dsconv_tc, args, what, "dsconv_tc", 2.0 * npix * cin * (k * k + cout), nbytes)
            return True
        want = self.fuse_ds == "1" or (self.fuse_ds == "auto" and s == 1 and x.h * x.w * cin <= 131072)
        if not (self.fuse_thin and want and self.lib.dbl_dsconv_supported(k, s, cin, cout)) or x.buf.data_ptr() == y.buf.data_ptr():
            return False
        wd = self._dev(wdw.reshape(cin, k * k).t(), torch.float32)
        bd = self._dev(bdw, torch.float32) if bdw is not None else None
        wp, bp = self._pack_mma(wpw), self._dev(bpw, torch.float32)
        npix = self.B * y.h * y.w
        args = (C.byref(x.desc), L.ptr(wd), L.ptr(bd), 1 if bdw is not None else 0, L.ptr(wp), L.ptr(bp),
                C.byref(r.desc) if r is not None else None, C.byref(y.desc), k, s, epi if r is not None else L.EPI_NONE)
        nbytes = 2 * self.B * x.h * x.w * cin + 2 * npix * cout * (2 if r is not None else 1) + 4 * k * k * cin + 2 * cin * cout
        self._emit(self.lib.dbl_dsconv, args, what, "dsconv", 2.0 * npix * cin * (k * k + cout), nbytes)
        return True

    def _dsconv(self, x: Act, p: str, y: Act, s: int = 1, epi=L.EPI_NONE, r=None):
        """DSConv (conv.py:91-108): depthwise -> pointwise -> BN -> SiLU, BN folded into the pointwise GEMM"""
        wdw = self.sd[p + ".dw.weight"].float()
        k = wdw.shape[-1]
        wf, bf = _fold(self.sd[p + ".pw.weight"], self.sd, p + ".bn")
        if self._dsconv_fused(x, wdw, None, wf, bf, y, s, epi, r, p):
            return
        t = self._alloc(x.h // s, x.w // s, x.c)
        self._dw(x, wdw, None, t, k, s, 1, L.ACT_NONE, p + ".dw")
        w, b = _fold(self.sd[p + ".pw.weight"], self.sd, p + ".bn")
        self._conv(t, w, b, y, 1, 1, L.ACT_SILU, epi, r, what=p + ".pw")

    # ------------------------------------------------------------------------------------------------ blocks
    def _lower_bottleneck(self, x: Act, p: str, y: Act, shortcut: bool):
        """Bottleneck.forward (block.py:344-357)"""
        c_ = self.sd[p + ".cv1.conv.weight"].shape[0]
        add = shortcut and x.c == y.c
        k33 = self.sd[p + ".cv1.conv.weight"].shape[-1] == 3 and self.sd[p + ".cv2.conv.weight"].shape[-1] == 3
        if self.fuse_thin and self.bn_tc and k33 and x.c == y.c and c_ * 2 == x.c and x.buf.data_ptr() != y.buf.data_ptr() \
                and self.lib.dbl_bottleneck_tc_supported(x.c, x.w, x.desc.pitch) and y.desc.pitch == y.c:
            # cv1 -> cv2 -> +x in one tcgen05 kernel over pixel-group rows with band-matrix filters (csrc/bottleneck_tc.cu)
            w1, b1 = _fold(self.sd[p + ".cv1.conv.weight"], self.sd, p + ".cv1.bn")
            w2, b2 = _fold(self.sd[p + ".cv2.conv.weight"], self.sd, p + ".cv2.bn")
            packed = torch.zeros((self.lib.dbl_bottleneck_tc_packed_bytes(x.c) + 15) // 16 * 4, dtype=torch.int32, device=self.dev)
            L.check(self.lib.dbl_bottleneck_tc_pack(L.ptr(self._pack_w(w1)), L.ptr(self._dev(b1, torch.float32)), L.ptr(self._pack_w(w2)),
                                                    L.ptr(self._dev(b2, torch.float32)), x.c, L.ptr(packed), L.current_stream()), "dbl_bottleneck_tc_pack")
            torch.cuda.synchronize(self.dev)
            self._keep.append(packed)
            npix = self.B * y.h * y.w
            self._emit(self.lib.dbl_bottleneck3x3_tc, (C.byref(x.desc), L.ptr(packed), 1 if add else 0, C.byref(y.desc)), p, "bottleneck_tc",
                       2.0 * npix * 9 * 2 * x.c * c_, 2 * npix * (x.c + y.c) + 2 * (w1.numel() + w2.numel()) + 4 * (c_ + y.c))
            return
        if self.fuse_thin and add and x.c in (16, 32, 64) and x.c <= self.fuse_bn_max_c and c_ * 2 == x.c and x.buf.data_ptr() != y.buf.data_ptr() \
                and self.sd[p + ".cv1.conv.weight"].shape[-1] == 3 and self.sd[p + ".cv2.conv.weight"].shape[-1] == 3:
            # cv1 -> cv2 -> +x in one kernel, hidden tensor in shared memory (csrc/conv_mma.cu)
            w1, b1 = _fold(self.sd[p + ".cv1.conv.weight"], self.sd, p + ".cv1.bn")
            w2, b2 = _fold(self.sd[p + ".cv2.conv.weight"], self.sd, p + ".cv2.bn")
            npix = self.B * y.h * y.w
            self._emit(self.lib.dbl_bottleneck3x3, (C.byref(x.desc), L.ptr(self._pack_mma(w1)), L.ptr(self._dev(b1, torch.float32)),
                                                    L.ptr(self._pack_mma(w2)), L.ptr(self._dev(b2, torch.float32)), 1, C.byref(y.desc)),
                       p, "bottleneck_mma", 2.0 * npix * 9 * 2 * x.c * c_, 2 * npix * (x.c + y.c) + 2 * (w1.numel() + w2.numel()) + 4 * (c_ + y.c))
            return
        t = self._alloc(x.h, x.w, c_)
        self._conv_bn(x, p + ".cv1", t)
        self._conv_bn(t, p + ".cv2", y, epi=L.EPI_ADD if add else L.EPI_NONE, r=x if add else None)

    def _lower_dsbottleneck_inplace(self, x: Act, p: str):
        """DSBottleneck.forward (block.py:1408-1444), residual variant, result written over x"""
        t = self._alloc(x.h, x.w, self.sd[p + ".cv1.pw.weight"].shape[0])
        self._dsconv(x, p + ".cv1", t)
        self._dsconv(t, p + ".cv2", x, epi=L.EPI_ADD, r=x)

    def _lower_dsc3k(self, x: Act, p: str, y: Act, n: int = 2):
        """DSC3k (block.py:1447-1503) over C3.forward (block.py:271-273)"""
        c_ = self.sd[p + ".cv1.conv.weight"].shape[0]
        t = self._alloc(x.h, x.w, 2 * c_)
        self._conv_bn_pair(x, p + ".cv1", p + ".cv2", t)
        for j in range(n):
            self._lower_dsbottleneck_inplace(t.slice(0, c_), f"{p}.m.{j}")
        self._conv_bn(t, p + ".cv3", y)

    def _lower_dsc3k2(self, x: Act, L_: LayerSpec, p: str, y: Act):
        """DSC3k2(dsc3k=True) (block.py:1505-1580) over C2f.forward (block.py:245-249)"""
        c = self.sd[p + ".cv1.conv.weight"].shape[0] // 2
        z = self._alloc(x.h, x.w, (2 + L_.n) * c)
        self._conv_bn(x, p + ".cv1", z.slice(0, 2 * c))
        for j in range(L_.n):
            src, dst = z.slice((1 + j) * c, c), z.slice((2 + j) * c, c)
            if L_.extra.get("dsc3k", True):
                self._lower_dsc3k(src, f"{p}.m.{j}", dst)
            else:  # plain DSBottleneck(c, c, e=1, k1=3, k2=7) (block.py:1565-1577): dst = src + cv2(cv1(src))
                t = self._alloc(x.h, x.w, c)
                self._dsconv(src, f"{p}.m.{j}.cv1", t)
                self._dsconv(t, f"{p}.m.{j}.cv2", dst, epi=L.EPI_ADD, r=src)
        self._conv_bn(z, p + ".cv2", y)

    def _lower_a2c2f(self, x: Act, L_: LayerSpec, p: str, y: Act):
        """A2C2f.forward (block.py:1392-1405), a2 = True: cv1 -> n x [ABlock, ABlock] -> cat -> cv2.  ABlock (block.py:1318-1322):
        x + AAttn(x), then x + mlp(x); AAttn (block.py:1236-1268): qkv 1x1 -> area attention (dbl_area_attention) -> + pe(v) (7x7 depthwise)
        -> proj 1x1.  Residual adds ride in the epilogues of proj and mlp.1."""
        c_ = self.sd[p + ".cv1.conv.weight"].shape[0]
        heads, area = c_ // 32, L_.extra["area"]
        z = self._alloc(x.h, x.w, (1 + L_.n) * c_)
        self._conv_bn(x, p + ".cv1", z.slice(0, c_))
        for j in range(L_.n):
            cur = z.slice(j * c_, c_)
            for b in range(2):
                q = f"{p}.m.{j}.{b}"
                out = z.slice((j + 1) * c_, c_) if b == 1 else self._alloc(x.h, x.w, c_)
                qkv, o, v, pe = (self._alloc(x.h, x.w, 3 * c_), self._alloc(x.h, x.w, c_), self._alloc(x.h, x.w, c_), self._alloc(x.h, x.w, c_))
                self._conv_bn(cur, q + ".attn.qkv", qkv, act=L.ACT_NONE)
                self._emit(self.lib.dbl_area_attention, (C.byref(qkv.desc), C.byref(o.desc), C.byref(v.desc), heads, area), q + ".attn.core")
                self._dw_bn(v, q + ".attn.pe", pe, L.ACT_NONE)
                self._emit(self.lib.dbl_gated_add, (C.byref(o.desc), C.byref(pe.desc), C.c_float(1.0), C.byref(o.desc), None), q + ".attn.pe_add")
                x1 = self._alloc(x.h, x.w, c_)
                self._conv_bn(o, q + ".attn.proj", x1, act=L.ACT_NONE, epi=L.EPI_ADD, r=cur)
                hdn = self._alloc(x.h, x.w, self.sd[q + ".mlp.0.conv.weight"].shape[0])
                self._conv_bn(x1, q + ".mlp.0", hdn)
                self._conv_bn(hdn, q + ".mlp.1", out, act=L.ACT_NONE, epi=L.EPI_ADD, r=x1)
                cur = out
        if L_.extra.get("residual"):  # l / x scales: x + gamma * cv2(...) with a per-channel learnable gamma (block.py:1403-1404)
            t = self._alloc(x.h, x.w, y.c)
            self._conv_bn(z, p + ".cv2", t)
            gm = self._dev(self.sd[p + ".gamma"], torch.float32)
            self._emit(self.lib.dbl_channel_scale_add, (C.byref(x.desc), C.byref(t.desc), L.ptr(gm), C.byref(y.desc)), p + ".gamma")
        else:
            self._conv_bn(z, p + ".cv2", y)

    def _dw_bn(self, x: Act, p: str, y: Act, act: int):
        """depthwise Conv module: Conv(c, c, k, g=c) + BN (+ SiLU), BN folded into the filter (GhostConv.cv2, conv.py:192)"""
        w, b = _fold(self.sd[p + ".conv.weight"], self.sd, p + ".bn")
        self._dw(x, w, b, y, w.shape[-1], 1, 1, act, p)

    def _lower_c3ghost(self, x: Act, L_: LayerSpec, p: str, y: Act):
        """C3Ghost (block.py:313-320) over C3.forward (block.py:271-273); m = n GhostBottleneck(c_, c_), stride 1 (block.py:323-341):
        x + GhostConv(act=False)(GhostConv(x)), GhostConv = [1x1 Conv | 5x5 depthwise Conv of it] (conv.py:184-197).  Both halves of a
        GhostConv are channel slices of one buffer, so its torch.cat costs nothing."""
        c_ = self.sd[p + ".cv1.conv.weight"].shape[0]
        t = self._alloc(x.h, x.w, 2 * c_)
        self._conv_bn_pair(x, p + ".cv1", p + ".cv2", t)
        a = t.slice(0, c_)
        for j in range(L_.n):
            q = f"{p}.m.{j}.conv"
            h = self.sd[q + ".0.cv1.conv.weight"].shape[0] * 2   # GhostBottleneck hidden width = c_ / 2
            g1, g2 = self._alloc(x.h, x.w, h), self._alloc(x.h, x.w, c_)
            self._conv_bn(a, q + ".0.cv1", g1.slice(0, h // 2))
            self._dw_bn(g1.slice(0, h // 2), q + ".0.cv2", g1.slice(h // 2, h // 2), L.ACT_SILU)
            self._conv_bn(g1, q + ".2.cv1", g2.slice(0, c_ // 2), act=L.ACT_NONE)
            self._dw_bn(g2.slice(0, c_ // 2), q + ".2.cv2", g2.slice(c_ // 2, c_ // 2), L.ACT_NONE)
            self._emit(self.lib.dbl_gated_add, (C.byref(a.desc), C.byref(g2.desc), C.c_float(1.0), C.byref(a.desc), None), f"{p}.m.{j}.add")
        self._conv_bn(t, p + ".cv3", y)

    def _lower_lsk(self, x: Act, p: str, y: Act):
        """LSKblock.forward (nn/modules_attention/LSKA.py:28-52)"""
        d, sd = x.c, self.sd
        a1, a2 = self._alloc(x.h, x.w, d), self._alloc(x.h, x.w, d)
        self._dw(x, sd[p + ".conv0.weight"].float(), sd[p + ".conv0.bias"].float(), a1, 5, 1, 1, L.ACT_NONE, p + ".conv0")
        self._dw(a1, sd[p + ".conv_spatial.weight"].float(), sd[p + ".conv_spatial.bias"].float(), a2, 7, 1, 3, L.ACT_NONE,
                 p + ".conv_spatial")
        t = self._alloc(x.h, x.w, d)
        self._conv_bias(a1, p + ".conv1", t.slice(0, d // 2))
        self._conv_bias(a2, p + ".conv2", t.slice(d // 2, d // 2))
        agg = self._alloc(x.h, x.w, 2, torch.float32)
        self._emit(self.lib.dbl_lsk_squeeze, (C.byref(t.desc), C.byref(agg.desc)), p + ".squeeze")
        g = self._alloc(x.h, x.w, d // 2)
        wsq = self._dev(sd[p + ".conv_squeeze.weight"].reshape(-1), torch.float32)
        bsq = self._dev(sd[p + ".conv_squeeze.bias"], torch.float32)
        self._emit(self.lib.dbl_lsk_gate, (C.byref(t.desc), C.byref(agg.desc), L.ptr(wsq), L.ptr(bsq), C.byref(g.desc)), p + ".gate")
        self._conv_bias(g, p + ".conv", y, epi=L.EPI_MUL, r=x)

    def _lower_c3ah(self, x: Act, p: str, y: Act, he: int):
        """C3AH.forward (block.py:1754-1795) with AdaHGComputation (block.py:1710-1752)"""
        c, sd = x.c, self.sd
        t = self._alloc(x.h, x.w, 2 * c)
        self._conv_bn_pair(x, p + ".cv1", p + ".cv2", t)
        g = p + ".m.hgnn"
        f32 = torch.float32
        wt = L.HypergraphWeights(
            L.ptr(self._dev(sd[g + ".edge_generator.prototype_base"], f32)),
            L.ptr(self._dev(sd[g + ".edge_generator.context_net.weight"], f32)),
            L.ptr(self._dev(sd[g + ".edge_generator.context_net.bias"], f32)),
            L.ptr(self._dev(sd[g + ".edge_generator.pre_head_proj.weight"], f32)),
            L.ptr(self._dev(sd[g + ".edge_generator.pre_head_proj.bias"], f32)),
            L.ptr(self._dev(sd[g + ".edge_proj.0.weight"], f32)), L.ptr(self._dev(sd[g + ".edge_proj.0.bias"], f32)),
            L.ptr(self._dev(sd[g + ".node_proj.0.weight"], f32)), L.ptr(self._dev(sd[g + ".node_proj.0.bias"], f32)),
            he, c // 16)
        nbytes = self.lib.dbl_hypergraph_workspace_bytes(self.B, x.h * x.w, c, he)
        ws = torch.zeros(nbytes // 4 + 4, dtype=f32, device=self.dev)
        self._keep += [wt, ws]
        tok = t.slice(0, c)
        self._emit(self.lib.dbl_hypergraph, (C.byref(tok.desc), C.byref(wt), C.byref(tok.desc), ws.data_ptr(), C.c_size_t(nbytes)),
                   p + ".hypergraph")
        self._conv_bn(t, p + ".cv3", y)

    def _lower_hyperace(self, xs: Sequence[Act], L_: LayerSpec, p: str, y: Act, f: Optional[Act] = None):
        """HyperACE.forward (block.py:1884-1895) incl. FuseModule.forward (block.py:1833-1840).  `f`: the fuse buffer when the P4 producer
        already wrote its slice there (no copy of the middle input)."""
        x0, x1, x2 = xs
        aliased = f is not None and x1.buf.data_ptr() == f.buf.data_ptr() and x1.off == x0.c
        if not aliased:
            f = self._alloc(x1.h, x1.w, x0.c + x1.c + x2.c)
        self._emit(self.lib.dbl_avgpool2, (C.byref(x0.desc), C.byref(f.slice(0, x0.c).desc)), p + ".fuse.down")
        if not aliased:
            self._emit(self.lib.dbl_copy, (C.byref(x1.desc), C.byref(f.slice(x0.c, x1.c).desc)), p + ".fuse.copy")
        self._emit(self.lib.dbl_upsample2_nearest, (C.byref(x2.desc), C.byref(f.slice(x0.c + x1.c, x2.c).desc)), p + ".fuse.up")
        x = self._alloc(x1.h, x1.w, x1.c)
        self._conv_bn(f, p + ".fuse.conv_out", x)
        c = self.sd[p + ".cv1.conv.weight"].shape[0] // 3
        n = L_.n
        z = self._alloc(x.h, x.w, (4 + n) * c)  # [y0 | out1 | y2 | m_0..m_{n-1} | out2]
        self._conv_bn(x, p + ".cv1", z.slice(0, 3 * c))
        self._lower_c3ah(z.slice(c, c), p + ".branch2", z.slice((3 + n) * c, c), L_.he)  # reads y1 before branch1 overwrites it
        for j in range(n):
            self._lower_dsc3k(z.slice((2 + j) * c, c), f"{p}.m.{j}", z.slice((3 + j) * c, c))
        self._lower_c3ah(z.slice(c, c), p + ".branch1", z.slice(c, c), L_.he)
        self._conv_bn(z, p + ".cv2", y)

    def _lower_dysample(self, x: Act, p: str, y: Act):
        """DySample.forward_lp + sample (nn/modules_upsample/DySample.py:49-68)"""
        off = self._alloc(x.h, x.w, 32, torch.float32)
        self._conv_bias(x, p + ".offset", off)
        self._emit(self.lib.dbl_dysample, (C.byref(x.desc), C.byref(off.desc), C.byref(y.desc), 4), p + ".sample")

    def _lower_detect(self, xs: Sequence[Act], p: str):
        """Detect.forward (nn/modules/head.py:108-118), legacy=False class branch (head.py:93-100)"""
        sd, nc = self.sd, self.spec.nc
        pitch = (64 + nc + 3) // 4 * 4
        self.head_pitch = pitch
        self.head: List[Act] = []
        c3_all = sd[f"{p}.cv3.0.0.1.conv.weight"].shape[0]
        c2_all = sd[f"{p}.cv2.0.0.conv.weight"].shape[0]
        fused = self.fuse_head and bool(self.lib.dbl_head_decode_supported(c2_all, c3_all, nc))
        self.head_fused = L.HeadFused() if fused else None
        for i, x in enumerate(xs):
            hd = self._alloc(x.h, x.w, pitch, torch.float32)
            self.head.append(hd)
            c2b = sd[f"{p}.cv2.{i}.0.conv.weight"].shape[0]
            b1, b2 = self._alloc(x.h, x.w, c2b), self._alloc(x.h, x.w, c2b)
            self._conv_bn(x, f"{p}.cv2.{i}.0", b1)
            self._conv_bn(b1, f"{p}.cv2.{i}.1", b2)
            if not fused:
                self._conv_bias(b2, f"{p}.cv2.{i}.2", hd.slice(0, 64))
            c3 = sd[f"{p}.cv3.{i}.0.1.conv.weight"].shape[0]
            cur = x
            for j, cj in ((0, x.c), (1, c3)):
                q = f"{p}.cv3.{i}.{j}"
                wdw, bdw = _fold(sd[q + ".0.conv.weight"], sd, q + ".0.bn")  # DWConv = Conv(g=c) + BN + SiLU
                nxt = self._alloc(x.h, x.w, c3)
                wpw, bpw = _fold(sd[q + ".1.conv.weight"], sd, q + ".1.bn")
                if not self._dsconv_fused(cur, wdw, bdw, wpw, bpw, nxt, 1, L.EPI_NONE, None, q):
                    d = self._alloc(x.h, x.w, cj)
                    self._dw(cur, wdw, bdw, d, 3, 1, 1, L.ACT_SILU, q + ".0")
                    self._conv_bn(d, q + ".1", nxt)
                cur = nxt
            if not fused:
                self._conv_bias(cur, f"{p}.cv3.{i}.2", hd.slice(64, nc))
                continue
            # fused tail: the two final 1x1 convs + decode + filter run in dbl_head_decode (launched after the forward list)
            w2 = sd[f"{p}.cv2.{i}.2.weight"].float()
            w3 = sd[f"{p}.cv3.{i}.2.weight"].float()
            ncp = (nc + 7) // 8 * 8
            w3p = torch.zeros(ncp, w3.shape[1], 1, 1)
            w3p[:nc] = w3
            lvl = self.head_fused.level[i]
            lvl.f2, lvl.f3 = b2.desc.data, cur.desc.data
            lvl.pitch2, lvl.pitch3 = b2.buf.shape[3], cur.buf.shape[3]
            lvl.w2_packed, lvl.w3_packed = L.ptr(self._pack_mma(w2)), L.ptr(self._pack_mma(w3p))
            lvl.b2 = L.ptr(self._dev(sd[f"{p}.cv2.{i}.2.bias"], torch.float32))
            lvl.b3 = L.ptr(self._dev(sd[f"{p}.cv3.{i}.2.bias"], torch.float32))
            lvl.logits = hd.buf.data_ptr()
            lvl.h, lvl.w = x.h, x.w
        if fused:
            hf = self.head_fused
            hf.num_levels, hf.batch, hf.nc, hf.c2, hf.c3, hf.logit_pitch = len(xs), self.B, nc, c2_all, c3_all, pitch

    # ------------------------------------------------------------------------------------------------ stem
    def _stem_pair_ok(self) -> bool:
        """layers 0 and 1 = Conv(3->c0, k3, s1) then Conv(c0->c1, k3, s2) with (c0,c1) the fused kernel is built for,
        and nobody else reads layer 0"""
        if not self.fuse_stem or len(self.spec.layers) < 3:
            return False
        l0, l1 = self.spec.layers[0], self.spec.layers[1]
        if not (l0.kind == "Conv" and l0.k == 3 and l0.s == 1 and l1.kind == "Conv" and l1.k == 3 and l1.s == 2):
            return False
        if (l0.c2, l1.c2) not in ((8, 16), (16, 32)) or list(l1.src) not in ([-1], [0]):
            return False
        return not any(0 in [j for j in Lx.src if j >= 0] for Lx in self.spec.layers[2:])

    def _lower_stem_pair(self, y: Act):
        """Conv(3->c0,s1) + Conv(c0->c1,s2) + the input cast/scale of BasePredictor.preprocess in one kernel"""
        sd = self.sd
        w0, b0 = _fold(sd["model.0.conv.weight"], sd, "model.0.bn")
        w1, b1 = _fold(sd["model.1.conv.weight"], sd, "model.1.bn")
        w0d, b0d, b1d = self._dev(w0, torch.float32), self._dev(b0, torch.float32), self._dev(b1, torch.float32)
        c0 = w0.shape[0]
        npix0, npix1 = self.B * self.H * self.W, self.B * y.h * y.w
        flops = 2.0 * 27 * c0 * npix0 + 2.0 * 9 * c0 * y.c * npix1
        if self.stem_tc and self.lib.dbl_stem_tc_supported(c0, y.c, self.H, self.W, y.desc.pitch):
            # tcgen05 form (csrc/stem_tc.cu): band matrices packed once per input type (the u8 blob carries the /255)
            w1d = self._pack_w(w1)
            blobs = []
            for u8 in (0, 1):
                blob = torch.zeros((self.lib.dbl_stem_tc_packed_bytes() + 15) // 16 * 4, dtype=torch.int32, device=self.dev)
                L.check(self.lib.dbl_stem_tc_pack(L.ptr(w0d), L.ptr(b0d), L.ptr(w1d), L.ptr(b1d), u8, L.ptr(blob), L.current_stream()), "dbl_stem_tc_pack")
                blobs.append(blob)
            torch.cuda.synchronize(self.dev)
            self._keep.extend(blobs)
            self._emit(self.lib.dbl_stem_pair_tc, (self.x_in.data_ptr(), 0, self.B, self.H, self.W, L.ptr(blobs[0]), C.byref(y.desc)),
                       "model.0+1.stem_pair", "stem_tc", flops, self.x_in.numel() * 4 + 2 * npix1 * y.c)
            self._stem_u8 = (self.x_in_u8.data_ptr(), 1, self.B, self.H, self.W, L.ptr(blobs[1]), C.byref(y.desc))
            self._stem_fn_u8 = self.lib.dbl_stem_pair_tc
            return
        w1p = self._pack_mma(w1)
        tail = (self.B, self.H, self.W, L.ptr(w0d), L.ptr(b0d), c0, L.ptr(w1p), L.ptr(b1d), C.byref(y.desc))
        self._emit(self.lib.dbl_stem_pair, (self.x_in.data_ptr(), 0) + tail, "model.0+1.stem_pair", "stem_pair", flops,
                   self.x_in.numel() * 4 + 2 * npix1 * y.c)
        self._stem_u8 = (self.x_in_u8.data_ptr(), 1) + tail
        self._stem_fn_u8 = self.lib.dbl_stem_pair

    # ------------------------------------------------------------------------------------------------ graph
    def _build(self):
        spec, sd = self.spec, self.sd
        # concat buffers are allocated first; producers that feed a Concat write straight into their slice
        alias: Dict[int, Tuple[int, int]] = {}
        for Lc in spec.layers:
            if Lc.kind == "Concat":
                off = 0
                for j, cj in zip(Lc.src, Lc.c1):
                    if j >= 0 and j not in alias and spec.layers[j].kind != "Concat":
                        alias[j] = (Lc.i, off)
                    off += cj
            elif Lc.kind == "HyperACE":  # FuseModule's cat([down(P3), P4, up(P5)]) (block.py:1833-1840): the P4 producer writes its slice directly
                j = Lc.src[1]
                if j >= 0 and j not in alias and spec.layers[j].kind != "Concat":
                    alias[j] = (Lc.i, Lc.extra["cins"][0])
        cat_buf: Dict[int, Act] = {}

        def out_act(Lx: LayerSpec) -> Act:
            h, w = self.H // Lx.extra["out_stride"], self.W // Lx.extra["out_stride"]
            if Lx.i in alias:
                ci, off = alias[Lx.i]
                if ci not in cat_buf:
                    tgt = spec.layers[ci]
                    cat_buf[ci] = self._alloc(h, w, sum(tgt.extra["cins"]) if tgt.kind == "HyperACE" else tgt.c2)
                return cat_buf[ci].slice(off, Lx.c2)
            if Lx.kind == "Concat":
                if Lx.i not in cat_buf:
                    cat_buf[Lx.i] = self._alloc(h, w, Lx.c2)
                return cat_buf[Lx.i]
            return self._alloc(h, w, Lx.c2)

        outs: List[Optional[Act]] = []
        self.layer_ops: List[Tuple[int, int]] = []  # [first, last) launch index of every top-level layer
        for Lx in spec.layers:
            p = f"model.{Lx.i}"
            ins = [outs[j] if j >= 0 else None for j in Lx.src]
            self.layer_ops.append((len(self.ops), len(self.ops)))
            if Lx.kind == "Detect":
                self._lower_detect(ins, p)
                outs.append(None)
                self.layer_ops[-1] = (self.layer_ops[-1][0], len(self.ops))
                continue
            if Lx.i == 0 and self._stem_pair_ok():
                outs.append(None)  # layer 0's output never exists: it is consumed inside the fused stem kernel
                continue
            y = out_act(Lx)
            if Lx.i == 1 and outs[0] is None:
                self._lower_stem_pair(y)
            elif Lx.i == 0:
                if not (Lx.kind == "Conv" and Lx.k == 3):
                    raise NotImplementedError("layer 0 must be a 3x3 Conv")
                w, b = _fold(sd[p + ".conv.weight"], sd, p + ".bn")
                wd, bd = self._dev(w, torch.float32), self._dev(b, torch.float32)
                self._emit(self.lib.dbl_stem_conv, (self.x_in.data_ptr(), self.B, self.H, self.W, L.ptr(wd), L.ptr(bd), C.byref(y.desc), Lx.s),
                           p + ".stem", "stem_conv", 2.0 * 27 * y.c * self.B * y.h * y.w)
                self._stem_u8 = (self.x_in_u8.data_ptr(), self.B, self.H, self.W, L.ptr(wd), L.ptr(bd), C.byref(y.desc), Lx.s)
                self._stem_fn_u8 = self.lib.dbl_stem_conv_u8
            elif Lx.kind == "Conv":
                g = Lx.extra.get("g", 1)
                if g > 1 or not self._conv_bn_thin(ins[0], p, y, Lx.s):
                    self._conv_bn(ins[0], p, y, s=Lx.s, g=g)
            elif Lx.kind == "DSConv":
                self._dsconv(ins[0], p, y, s=Lx.s)
            elif Lx.kind == "Bottleneck":
                cur = ins[0]
                for r in range(Lx.n):
                    dst = y if r == Lx.n - 1 else self._alloc(y.h, y.w, Lx.c2)
                    self._lower_bottleneck(cur, f"{p}.{r}" if Lx.n > 1 else p, dst, Lx.shortcut)
                    cur = dst
            elif Lx.kind == "DSC3k2":
                self._lower_dsc3k2(ins[0], Lx, p, y)
            elif Lx.kind == "C3Ghost":
                self._lower_c3ghost(ins[0], Lx, p, y)
            elif Lx.kind == "A2C2f":
                self._lower_a2c2f(ins[0], Lx, p, y)
            elif Lx.kind == "nn.Upsample":
                self._emit(self.lib.dbl_upsample2_nearest, (C.byref(ins[0].desc), C.byref(y.desc)), p + ".up")
            elif Lx.kind == "LSKblock":
                cur = ins[0]
                for r in range(Lx.n):
                    dst = y if r == Lx.n - 1 else self._alloc(y.h, y.w, Lx.c2)
                    self._lower_lsk(cur, f"{p}.{r}" if Lx.n > 1 else p, dst)
                    cur = dst
            elif Lx.kind == "HyperACE":
                self._lower_hyperace(ins, Lx, p, y, cat_buf.get(Lx.i))
            elif Lx.kind == "DySample":
                self._lower_dysample(ins[0], p, y)
            elif Lx.kind == "DownsampleConv":
                t = self._alloc(y.h, y.w, ins[0].c)
                self._emit(self.lib.dbl_avgpool2, (C.byref(ins[0].desc), C.byref(t.desc)), p + ".down")
                self._conv_bn(t, p + ".channel_adjust", y)
            elif Lx.kind == "FullPAD_Tunnel":
                gate = float(sd[p + ".gate"])
                self._emit(self.lib.dbl_gated_add, (C.byref(ins[0].desc), C.byref(ins[1].desc), C.c_float(gate), C.byref(y.desc), None),
                           p + ".gate")
            elif Lx.kind == "Concat":
                off = 0
                for j, a in zip(Lx.src, ins):
                    if alias.get(j, (None,))[0] != Lx.i:
                        self._emit(self.lib.dbl_copy, (C.byref(a.desc), C.byref(y.slice(off, a.c).desc)), f"{p}.copy{j}")
                    off += a.c
            else:
                raise NotImplementedError(Lx.kind)
            outs.append(y)
            self.layer_ops[-1] = (self.layer_ops[-1][0], len(self.ops))
        self.layer_out = outs
        det = spec.detect
        self.level_hw = [(a.h, a.w) for a in self.head]
        self.A = sum(h * w for h, w in self.level_hw)
        lv = L.HeadLevels()
        for i, a in enumerate(self.head):
            lv.data[i], lv.h[i], lv.w[i], lv.stride[i] = a.buf.data_ptr(), a.h, a.w, float(det.extra["strides"][i])
        lv.num_levels, lv.batch, lv.nc, lv.pitch = len(self.head), self.B, spec.nc, self.head_pitch
        self.levels = lv
        if self.head_fused is not None:  # two descriptions of the fused tail: with logits (forward) and without (detect)
            for i in range(len(self.head)):
                self.head_fused.level[i].stride = float(det.extra["strides"][i])
            self.head_fused_nolog = L.HeadFused.from_buffer_copy(self.head_fused)
            for i in range(len(self.head)):
                self.head_fused_nolog.level[i].logits = None
        self.y = torch.zeros(self.B, 4 + spec.nc, self.A, dtype=torch.float32, device=self.dev)

    # ------------------------------------------------------------------------------------------------ lanes
    @staticmethod
    def _region(t: "L.Tensor"):
        es = 2 if t.dtype == L.DBL_BF16 else 4
        stride = t.pitch * es
        return (t.data, t.data + ((t.n * t.h * t.w - 1) * t.pitch + t.c) * es, stride, t.c * es)

    @staticmethod
    def _overlap(a, b) -> bool:
        if a[1] <= b[0] or b[1] <= a[0]:
            return False
        if a[2] != b[2]:
            return True  # different pixel pitch over the same memory: be conservative
        d = (b[0] - a[0]) % a[2]
        return d < a[3] or d + b[3] > a[2]

    def _plan_lanes(self):
        """Dependency DAG of the launch list (RAW, WAW and WAR hazards on the channel-sliced buffers) and a greedy
        assignment of every launch to one of `n_lanes` streams.  self.plan[i] = (lane, [ops whose completion event this
        launch must wait for because they run on another lane])."""
        n = len(self.ops)
        deps: List[set] = [set() for _ in range(n)]
        anc = [0] * n  # ancestor bitsets
        for j in range(n):
            rj, wj = self.op_rw[j]
            for i in range(j - 1, -1, -1):
                if (anc[j] >> i) & 1:
                    continue  # already ordered through another dependency
                ri, wi = self.op_rw[i]
                if any(self._overlap(w, r) for w in wi for r in rj) or any(self._overlap(w, w2) for w in wi for w2 in wj) \
                        or any(self._overlap(r, w) for r in ri for w in wj):
                    deps[j].add(i)
                    anc[j] |= anc[i] | (1 << i)
        lane_tail = [-1] * self.n_lanes
        self.plan: List[Tuple[int, List[int]]] = []
        lane_of = [0] * n
        for j in range(n):
            lane = None
            if self.n_lanes > 1:
                for i in sorted(deps[j], reverse=True):  # continue the chain of the most recent predecessor if it is a tail
                    if lane_tail[lane_of[i]] == i:
                        lane = lane_of[i]
                        break
                if lane is None:  # a lane whose tail is an ancestor anyway (no false dependency), else the least recently used
                    cands = [l for l in range(self.n_lanes) if lane_tail[l] < 0 or (anc[j] >> lane_tail[l]) & 1]
                    lane = cands[0] if cands else min(range(self.n_lanes), key=lambda l: lane_tail[l])
            else:
                lane = 0
            if j == 0:
                lane = 0
            lane_of[j] = lane
            lane_tail[lane] = j
            self.plan.append((lane, sorted(i for i in deps[j] if lane_of[i] != lane)))
        self._signals = set(i for _, w in self.plan for i in w)  # launches that need a completion event
        self._side_streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.n_lanes - 1)]

    # ------------------------------------------------------------------------------------------------ execution
    def _launch_forward(self, stream: int, lanes: bool = False):
        if lanes and self.n_lanes > 1:
            return self._launch_forward_lanes()
        for i, (fn, args, what) in enumerate(self.ops):
            if i == 0 and self._u8:  # uint8 input: same stem, the /255 happens in its loader
                fn, args = self._stem_fn_u8, self._stem_u8
            rc = fn(*args, stream)
            if rc:
                raise L.DblError(f"{what}: code {rc}: {self.lib.dbl_last_error().decode()}")

    def _launch_forward_lanes(self):
        """Same launches, spread over the planned lanes (lane 0 = the current stream).  Cross-lane dependencies are CUDA
        events; under stream capture they become edges of the graph.  All lanes are joined back before returning."""
        main = torch.cuda.current_stream(self.dev)
        streams = [main] + self._side_streams
        started = [True] + [False] * (self.n_lanes - 1)
        fork = torch.cuda.Event()
        fork.record(main)
        done: Dict[int, torch.cuda.Event] = {}
        for i, (fn, args, what) in enumerate(self.ops):
            if i == 0 and self._u8:
                fn, args = self._stem_fn_u8, self._stem_u8
            lane, waits = self.plan[i]
            st = streams[lane]
            if not started[lane]:
                st.wait_event(fork)
                started[lane] = True
            for d in waits:
                st.wait_event(done[d])
            rc = fn(*args, st.cuda_stream)
            if rc:
                raise L.DblError(f"{what}: code {rc}: {self.lib.dbl_last_error().decode()}")
            if i in self._signals:
                ev = torch.cuda.Event()
                ev.record(st)
                done[i] = ev
        for lane in range(1, self.n_lanes):
            if started[lane]:
                main.wait_stream(streams[lane])

    def _launch_decode(self, stream: int):
        if self.head_fused is not None:  # final head convs + decode in one kernel; also materialises the raw logits (feats)
            L.check(self.lib.dbl_head_decode(C.byref(self.head_fused), self.y.data_ptr(), C.c_float(0.0), 0, None, 0, None, stream),
                    "dbl_head_decode")
            return
        L.check(self.lib.dbl_detect_decode(C.byref(self.levels), self.y.data_ptr(), stream), "dbl_detect_decode")

    def _post(self, key: tuple) -> dict:
        if key not in self._post_bufs:
            conf, iou, agnostic, multi_label, max_det, classes, max_nms, max_wh, scaled, labels_cap = key
            ml = bool(multi_label) and self.spec.nc > 1
            cap = (self.A + labels_cap) * (self.spec.nc if ml else 1)  # a-priori labels (validator path) are appended to the candidates
            dev, B = self.dev, self.B
            d = dict(
                box=torch.zeros(B, cap, 4, device=dev), score=torch.zeros(B, cap, device=dev),
                cls=torch.zeros(B, cap, dtype=torch.int32, device=dev), pos=torch.zeros(B, cap, dtype=torch.int32, device=dev),
                count=torch.zeros(B, dtype=torch.int32, device=dev),
                keys=torch.zeros(B * self.lib.dbl_nms_sort_capacity(cap), dtype=torch.int64, device=dev),
                rows=torch.zeros(B, max_det, 6, device=dev), out_pos=torch.zeros(B, max_det, dtype=torch.int32, device=dev),
                out_count=torch.zeros(B, dtype=torch.int32, device=dev),
                classes=torch.tensor(list(classes), dtype=torch.int32, device=dev) if classes else None,
                # per-image (pad_x, pad_y, gain, w0, h0) of scale_boxes / clip_boxes, read by the NMS epilogue; identity until set_scale()
                scale=torch.tensor([[0.0, 0.0, 1.0, float(self.W), float(self.H)]] * B, dtype=torch.float32, device=dev) if scaled else None,
                labels=torch.zeros(max(1, B * labels_cap), 5, device=dev) if labels_cap else None,
                label_offs=torch.zeros(B + 1, dtype=torch.int32, device=dev) if labels_cap else None,
            )
            d["cand"] = L.Candidates(d["box"].data_ptr(), d["score"].data_ptr(), d["cls"].data_ptr(), d["pos"].data_ptr(),
                                     d["count"].data_ptr(), cap, B)
            self._post_bufs[key] = d
        return self._post_bufs[key]

    def _launch_filter(self, key: tuple, d: dict, stream: int):
        conf, iou, agnostic, multi_label, max_det, classes, max_nms, max_wh, scaled, labels_cap = key
        ncls = len(classes) if classes else 0
        if self.head_fused is not None:
            rc = self.lib.dbl_head_decode(C.byref(self.head_fused_nolog), None, C.c_float(conf), int(bool(multi_label)), L.ptr(d["classes"]),
                                          ncls, C.byref(d["cand"]), stream)
            L.check(rc, "dbl_head_decode")
        else:
            rc = self.lib.dbl_decode_filter(C.byref(self.levels), C.c_float(conf), int(bool(multi_label)), L.ptr(d["classes"]), ncls,
                                            C.byref(d["cand"]), stream)
            L.check(rc, "dbl_decode_filter")
        if labels_cap:  # validator path: non_max_suppression(labels=...) (utils/ops.py:255-261)
            L.check(self.lib.dbl_nms_append_labels(d["labels"].data_ptr(), d["label_offs"].data_ptr(), labels_cap, self.A, self.spec.nc,
                                                   int(bool(multi_label)), C.c_float(conf), L.ptr(d["classes"]), ncls, C.byref(d["cand"]), stream),
                    "dbl_nms_append_labels")

    def _launch_nms(self, key: tuple, d: dict, stream: int):
        conf, iou, agnostic, multi_label, max_det, classes, max_nms, max_wh, scaled, labels_cap = key
        L.check(self.lib.dbl_nms_scaled(C.byref(d["cand"]), C.c_float(iou), int(bool(agnostic)), C.c_float(max_wh), int(max_nms), int(max_det),
                                        d["keys"].data_ptr(), L.ptr(d["scale"]), d["rows"].data_ptr(), d["out_pos"].data_ptr(),
                                        d["out_count"].data_ptr(), stream), "dbl_nms")

    def _launch_post(self, key: tuple, stream: int):
        d = self._post(key)
        self._launch_filter(key, d, stream)
        self._launch_nms(key, d, stream)

    def _run(self, key: Optional[tuple], want_y: bool):
        """Launch forward (+decode) (+post-process) on the current stream, through a cached CUDA graph when enabled."""
        gkey = (key, want_y, self._u8)
        if not self.use_graph:
            s = L.current_stream()
            self._launch_forward(s)
            if want_y:
                self._launch_decode(s)
            if key is not None:
                self._launch_post(key, s)
            return
        g = self._graphs.get(gkey)
        if g is None:
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            n0 = self.lib.dbl_launch_count()
            with torch.cuda.stream(s):  # warm-up outside capture: one-time attribute / descriptor setup happens here
                self._launch_forward(s.cuda_stream, lanes=True)
                if want_y:
                    self._launch_decode(s.cuda_stream)
                if key is not None:
                    self._launch_post(key, s.cuda_stream)
            s.synchronize()
            self.kernels_per_step[gkey] = int(self.lib.dbl_launch_count() - n0)  # kernels one replay of this graph runs
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                cs = L.current_stream()
                self._launch_forward(cs, lanes=True)
                if want_y:
                    self._launch_decode(cs)
                if key is not None:
                    self._launch_post(key, cs)
            self._graphs[gkey] = g
        g.replay()

    @staticmethod
    def post_key(conf=0.25, iou=0.45, agnostic=False, multi_label=False, max_det=300, classes=None, max_nms=30000, max_wh=7680,
                 scaled=False, labels_cap=0):
        """scaled: the NMS epilogue applies scale_boxes + clip_boxes with the per-image records of set_scale();
        labels_cap > 0: up to that many a-priori labels per image (set_labels()) join the candidates (validator path)."""
        return (float(conf), float(iou), bool(agnostic), bool(multi_label), int(max_det), tuple(classes) if classes else (),
                int(max_nms), float(max_wh), bool(scaled), int(labels_cap))

    def set_scale(self, key: tuple, records: torch.Tensor):
        """records [B,5] f32 = (pad_x, pad_y, gain, w0, h0) per image (host or device); stream-ordered copy into the table the
        captured NMS launch reads."""
        self._post(key)["scale"].copy_(records.to(torch.float32).reshape(self.B, 5), non_blocking=True)

    def set_labels(self, key: tuple, labels: Sequence[torch.Tensor]):
        """labels: per image an (n_i, 5) tensor (cls, x, y, w, h) in network-input pixels -- DetectionValidator.lb
        (models/yolo/detect/val.py:60-65).  Packed into the fixed-capacity device buffers the captured launch reads."""
        d, cap = self._post(key), key[9]
        if len(labels) != self.B:
            raise ValueError(f"expected {self.B} label tensors")
        offs, parts = [0], []
        for lb in labels:
            lb = torch.as_tensor(lb, dtype=torch.float32).reshape(-1, 5)
            if lb.shape[0] > cap:
                raise ValueError(f"{lb.shape[0]} labels in one image exceed labels_cap={cap}")
            parts.append(lb.cpu())
            offs.append(offs[-1] + lb.shape[0])
        if offs[-1]:
            d["labels"][: offs[-1]].copy_(torch.cat(parts), non_blocking=False)
        d["label_offs"].copy_(torch.tensor(offs, dtype=torch.int32), non_blocking=False)

    def set_input(self, im: torch.Tensor):
        """Stage one batch: float BCHW in 0..1 (LoadTensor sources) or uint8 BCHW in 0..255 (image sources); `im` may
        live in (pinned) host memory, in which case this is the step's host->device copy."""
        if tuple(im.shape) != tuple(self.x_in.shape):
            raise ValueError(f"engine was built for input {tuple(self.x_in.shape)}, got {tuple(im.shape)}")
        self._u8 = im.dtype == torch.uint8
        (self.x_in_u8 if self._u8 else self.x_in).copy_(im, non_blocking=True)

    def forward(self, im: torch.Tensor):
        """(y [B,4+nc,A] fp32, feats list of [B,64+nc,h,w] fp32 views) like DetectionModel.forward in eval mode."""
        with torch.cuda.device(self.dev):
            self.set_input(im)
            self._run(None, True)
        feats = [a.buf[..., : 64 + self.spec.nc].permute(0, 3, 1, 2) for a in self.head]
        return self.y, feats

    def detect(self, im: Optional[torch.Tensor], **nms_kw):
        """Fused path: image batch -> (rows [B,max_det,6], counts [B], kept anchor positions [B,max_det]); no host sync."""
        key = self.post_key(**nms_kw)
        with torch.cuda.device(self.dev):
            if im is not None:
                self.set_input(im)
            self._run(key, False)
        d = self._post(key)
        return d["rows"], d["out_count"], d["out_pos"]

    def profile_ops(self, reps: int = 5, post_key: Optional[tuple] = None) -> List[dict]:
        """Time every launch of one step on its own (CUDA events on the launching stream, eager, after one warm-up pass).

        Returns op_meta rows extended with ``ms`` (mean over reps).  Used by bench.py for the per-kernel roofline table;
        the whole-step number always comes from graph replays, never from this."""
        rows = []
        with torch.cuda.device(self.dev):
            st = L.current_stream()
            self._launch_forward(st)
            if post_key is not None:
                self._launch_post(post_key, st)
            torch.cuda.synchronize(self.dev)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(len(self.ops) + 2)]
            acc = [0.0] * (len(self.ops) + 2)
            for _ in range(reps):
                for i, (fn, args, what) in enumerate(self.ops):
                    ev[i][0].record()
                    fn(*args, st)
                    ev[i][1].record()
                if post_key is not None:
                    d = self._post(post_key)
                    n = len(self.ops)
                    ev[n][0].record()
                    self._launch_filter(post_key, d, st)
                    ev[n][1].record()
                    ev[n + 1][0].record()
                    self._launch_nms(post_key, d, st)
                    ev[n + 1][1].record()
                torch.cuda.synchronize(self.dev)
                for i in range(len(acc)):
                    if i < len(self.ops) or post_key is not None:
                        acc[i] += ev[i][0].elapsed_time(ev[i][1])
            for i, m in enumerate(self.op_meta):
                rows.append(dict(m, ms=acc[i] / reps))
            if post_key is not None:
                no = 64 + self.spec.nc
                if self.head_fused is not None:
                    c2, c3 = self.head_fused.c2, self.head_fused.c3
                    rows.append(dict(what="head_decode", kern="head_decode", flops=2.0 * self.B * self.A * (c2 * 64 + c3 * self.spec.nc),
                                     bytes=float(self.B * self.A * (c2 + c3) * 2), ms=acc[len(self.ops)] / reps))
                else:
                    rows.append(dict(what="decode_filter", kern="decode_filter", flops=0.0, bytes=float(self.B * self.A * no * 4),
                                     ms=acc[len(self.ops)] / reps))
                rows.append(dict(what="sort_nms", kern="sort_nms", flops=0.0, bytes=0.0, ms=acc[len(self.ops) + 1] / reps))
        return rows

    def debug_layers_forced(self, im: torch.Tensor, ref_outs: Sequence[torch.Tensor]) -> List[Optional[torch.Tensor]]:
        """Teacher-forced run (parity tests): before each top-level layer executes, its inputs are overwritten with the
        given reference outputs (fp32 NCHW, rounded to bf16 here), so errors cannot accumulate from layer to layer.
        Returns every layer's own output; the last entry is the stacked raw head logits (B, 64+nc, A)."""
        res: List[Optional[torch.Tensor]] = []
        with torch.cuda.device(self.dev):
            self.set_input(im)
            st = L.current_stream()
            for Lx, (a, b) in zip(self.spec.layers, self.layer_ops):
                for j in Lx.src:
                    if j >= 0:
                        act = self.layer_out[j]
                        act.buf[..., act.off:act.off + act.c] = ref_outs[j].to(self.dev).permute(0, 2, 3, 1).to(act.buf.dtype)
                for fn, args, what in self.ops[a:b]:
                    rc = fn(*args, st)
                    if rc:
                        raise L.DblError(f"{what}: code {rc}: {self.lib.dbl_last_error().decode()}")
                torch.cuda.synchronize(self.dev)
                if Lx.kind == "Detect":
                    no = 64 + self.spec.nc
                    res.append(torch.cat([h.buf[..., :no].permute(0, 3, 1, 2).reshape(self.B, no, -1) for h in self.head], 2).cpu())
                else:
                    res.append(self.layer_out[Lx.i].nchw())
        return res

    def debug_layers(self, im: torch.Tensor) -> List[Optional[torch.Tensor]]:
        """fp32 NCHW copy of every top-level layer output (parity tests)."""
        with torch.cuda.device(self.dev):
            self.set_input(im)
            self._launch_forward(L.current_stream())
            torch.cuda.synchronize(self.dev)
            return [a.nchw() if a is not None else None for a in self.layer_out]
