"""Per-kernel parity (-m gpu): each C-ABI entry point vs the CPU oracle / plain torch fp32 on identical bf16-rounded inputs.

Tolerances: bf16 outputs carry one rounding (2^-9 relative) on top of fp32 accumulation-order noise -> rel 2^-7 + 2e-2
absolute on O(1..10) activations; fp32 outputs 1e-3 relative.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import dbl_oracle as O
from yolo_dbl_b200 import _lib as L

from gpu_util import assert_close_bf16, bf16_round, conv2d, dev, empty_nhwc, from_nhwc, pack_w, stream, to_nhwc

pytestmark = pytest.mark.gpu


def _act_ref(v, act):
    return F.silu(v) if act == L.ACT_SILU else (F.gelu(v) if act == L.ACT_GELU else v)


CONV_CASES = [
    # n, h, w, cin, cout, k, s, act, epi, ydtype, xpad, ypad
    (2, 20, 20, 64, 64, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (2, 20, 20, 64, 32, 3, 1, L.ACT_SILU, L.EPI_ADD, torch.bfloat16, 0, 0),
    (1, 40, 40, 32, 16, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 24, 36, 16, 8, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (2, 40, 40, 16, 32, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 80, 80, 64, 64, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (2, 40, 40, 384, 64, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 20, 20, 128, 128, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 64, 256),
    (2, 40, 40, 64, 128, 1, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (3, 20, 20, 384, 256, 1, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (2, 20, 20, 128, 256, 1, 1, L.ACT_NONE, L.EPI_MUL, torch.bfloat16, 0, 0),
    (2, 20, 20, 64, 64, 1, 1, L.ACT_GELU, L.EPI_ADD, torch.bfloat16, 64, 64),
    (2, 40, 40, 64, 3, 1, 1, L.ACT_NONE, L.EPI_NONE, torch.float32, 0, 65),
    (1, 40, 40, 80, 80, 1, 1, L.ACT_NONE, L.EPI_NONE, torch.float32, 0, 64),
    (2, 40, 40, 128, 32, 1, 1, L.ACT_NONE, L.EPI_NONE, torch.float32, 0, 0),
    (1, 20, 20, 256, 320, 1, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 33, 50, 32, 48, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    # TMA-store epilogue of the 1x1 GEMM: partial last 64-channel sub-tile (80 = 64 + 16) into a channel slice, partial last pixel tile
    (1, 40, 40, 64, 80, 1, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 48),
    (2, 33, 31, 32, 24, 1, 1, L.ACT_SILU, L.EPI_ADD, torch.bfloat16, 0, 8),
    (3, 40, 40, 128, 192, 1, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 64, 64),
]


def _run_conv_case(case, impl, seed=0):
    n, h, w, cin, cout, k, s, act, epi, ydt, xpad, ypad = case
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wt = bf16_round(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.5
    ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
    r = bf16_round(torch.randn(n, cout, ho, wo, generator=g)) if epi else None
    ref = _act_ref(F.conv2d(x, wt, b, s, k // 2), act)
    if epi == L.EPI_ADD:
        ref = ref + r
    elif epi == L.EPI_MUL:
        ref = ref * r
    xbuf, xv = to_nhwc(x, pad_c=xpad, c_off=xpad // 2 // 8 * 8)
    ybuf, yv = empty_nhwc(n, ho, wo, cout, ydt, pad_c=ypad, c_off=(ypad // 2 // 8 * 8) if ydt == torch.bfloat16 else (64 if ypad >= 64 else 0))
    rv = None
    if epi:
        rbuf, rv = to_nhwc(r)
    rc = conv2d(xv, pack_w(wt), b.to(dev()), rv, yv, k, s, act, epi, impl)
    assert rc == 0, L.load().dbl_last_error().decode()
    off = yv.data - ybuf.data_ptr()
    got = from_nhwc(ybuf, off // ybuf.element_size(), cout)
    if ydt == torch.float32:
        torch.testing.assert_close(got, ref, rtol=1e-3, atol=2e-3)
    else:
        assert_close_bf16(got, ref, f"conv {case} impl={impl}")
    # channels outside the view must be untouched
    if ypad:
        mask = torch.ones(ybuf.shape[-1], dtype=torch.bool)
        c0 = off // ybuf.element_size()
        mask[c0:c0 + cout] = False
        assert bool((ybuf[..., mask.to(dev())].float() == -3.0).all()), "kernel wrote outside its channel slice"


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_direct(case):
    _run_conv_case(case, L.IMPL_DIRECT)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_tcgen05(case):
    _run_conv_case(case, L.IMPL_TCGEN05)


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[5] == 3])
def test_conv2d_tcgen05_tma_generation(case):
    _run_conv_case(case, L.IMPL_TCGEN05_TMA)


THIN_CASES = [  # the 8-channel path (two filter taps per K=16 step) and awkward tile geometries of the halo-reuse kernel
    (2, 64, 96, 8, 16, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (2, 48, 80, 8, 16, 3, 1, L.ACT_SILU, L.EPI_ADD, torch.bfloat16, 0, 0),
    (1, 320, 320, 16, 8, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 20, 20, 256, 64, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 128, 0),
    (1, 40, 40, 192, 64, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (2, 20, 20, 128, 320, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 18, 130, 32, 32, 3, 1, L.ACT_NONE, L.EPI_NONE, torch.float32, 0, 0),
    (3, 34, 38, 48, 24, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
]


@pytest.mark.parametrize("case", THIN_CASES)
def test_conv3x3_halo_reuse_kernel(case):
    _run_conv_case(case, L.IMPL_TCGEN05, seed=7)
    _run_conv_case(case, L.IMPL_DIRECT, seed=7)


S2_TC_CASES = [  # thin stride-2 down-samplers on the tcgen05 pixel-pair kernel (conv_s2_tc.cu), reached through dbl_conv2d(AUTO)
    (2, 40, 40, 16, 32, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 64, 96, 16, 32, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 32),   # output into a channel slice
    (3, 34, 38, 16, 32, 3, 2, L.ACT_NONE, L.EPI_NONE, torch.bfloat16, 0, 0),    # partial tiles in both directions
    (2, 40, 40, 32, 64, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 50, 130, 32, 64, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 64),
    (4, 160, 160, 16, 32, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),  # more tiles than SMs: every CTA loops
    (8, 160, 160, 32, 64, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 8, 8, 16, 32, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
]


@pytest.mark.parametrize("case", S2_TC_CASES)
def test_conv3x3_stride2_pixel_pair_kernel(case):
    n, h, w, cin, cout = case[:5]
    lib = L.require_device()
    xt = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=dev())
    yt = torch.empty(n, h // 2, w // 2, cout, dtype=torch.bfloat16, device=dev())
    xv, yv = L.view(xt), L.view(yt)
    assert lib.dbl_conv2d_which(C.byref(xv), C.byref(yv), 3, 2, 1, 1, L.IMPL_AUTO) == 5
    _run_conv_case(case, L.IMPL_AUTO, seed=11)


T5_CASES = [  # stride-1 3x3 with the whole reduction depth in one tile (conv_tc5.cu), reached through dbl_conv2d(AUTO)
    (2, 20, 20, 64, 32, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (2, 20, 20, 32, 64, 3, 1, L.ACT_SILU, L.EPI_ADD, torch.bfloat16, 0, 0),
    (1, 33, 50, 32, 64, 3, 1, L.ACT_NONE, L.EPI_MUL, torch.bfloat16, 32, 64),   # channel slices on both sides, partial tiles
    (1, 40, 40, 64, 64, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 64, 0),
    (2, 17, 23, 64, 64, 3, 1, L.ACT_SILU, L.EPI_ADD, torch.bfloat16, 0, 0),
    (1, 24, 36, 32, 32, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 40, 72, 32, 16, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 40, 72, 16, 32, 3, 1, L.ACT_SILU, L.EPI_ADD, torch.bfloat16, 0, 0),
    (1, 30, 30, 16, 16, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 16),
    (8, 80, 80, 64, 32, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),    # more tiles than SMs: every CTA loops
    (8, 80, 80, 32, 64, 3, 1, L.ACT_SILU, L.EPI_ADD, torch.bfloat16, 0, 0),
    (6, 80, 80, 64, 64, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 4, 4, 64, 32, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
]


@pytest.mark.parametrize("case", T5_CASES)
def test_conv3x3_single_chunk_kernel(case):
    n, h, w, cin, cout = case[:5]
    lib = L.require_device()
    xt = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=dev())
    yt = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev())
    xv, yv = L.view(xt), L.view(yt)
    assert lib.dbl_conv2d_which(C.byref(xv), C.byref(yv), 3, 1, 1, 1, L.IMPL_AUTO) == 6
    _run_conv_case(case, L.IMPL_AUTO, seed=13)


def test_conv2d_tcgen05_many_tiles_persistent():
    # more tiles than SMs so every CTA loops (pipeline phases, TMEM double buffering)
    _run_conv_case((8, 80, 80, 64, 64, 3, 1, L.ACT_SILU, L.EPI_ADD, torch.bfloat16, 0, 0), L.IMPL_TCGEN05, seed=3)
    _run_conv_case((16, 80, 80, 64, 128, 1, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0), L.IMPL_TCGEN05, seed=4)
    _run_conv_case((4, 160, 160, 16, 32, 3, 2, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0), L.IMPL_TCGEN05, seed=5)


def test_stem_conv():
    g = torch.Generator().manual_seed(1)
    for cout in (8, 16, 32):
        x = torch.rand(2, 3, 40, 56, generator=g)
        wt = torch.randn(cout, 3, 3, 3, generator=g) / 27 ** 0.5
        b = torch.randn(cout, generator=g) * 0.3
        ybuf, yv = empty_nhwc(2, 40, 56, cout)
        xd, wd, bd = x.to(dev()), wt.to(dev()), b.to(dev())  # keep the device copies alive across the launch
        rc = L.require_device().dbl_stem_conv(xd.data_ptr(), 2, 40, 56, wd.data_ptr(), bd.data_ptr(), C.byref(yv), 1, stream())
        torch.cuda.synchronize()
        assert rc == 0
        assert_close_bf16(from_nhwc(ybuf, 0, cout), F.silu(F.conv2d(x, wt, b, 1, 1)), f"stem cout={cout}")


@pytest.mark.parametrize("k,s,d,act,bias", [(3, 1, 1, L.ACT_NONE, False), (3, 2, 1, L.ACT_NONE, False), (7, 1, 1, L.ACT_NONE, False),
                                            (5, 1, 1, L.ACT_NONE, True), (7, 1, 3, L.ACT_NONE, True), (3, 1, 1, L.ACT_SILU, True)])
@pytest.mark.parametrize("geom", [(2, 64, 20, 28), (1, 24, 10, 30), (3, 128, 6, 6)])
def test_dwconv(k, s, d, act, bias, geom):
    g = torch.Generator().manual_seed(2)
    n, c, h, w = geom
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    wt = torch.randn(c, 1, k, k, generator=g) / k
    b = torch.randn(c, generator=g) if bias else None
    pad = d * (k - 1) // 2
    ref = _act_ref(F.conv2d(x, wt, b, s, pad, d, c), act)
    xbuf, xv = to_nhwc(x, pad_c=8)
    ybuf, yv = empty_nhwc(n, ref.shape[2], ref.shape[3], c)
    wd = wt.reshape(c, k * k).t().contiguous().to(dev())
    bd = None if b is None else b.to(dev())
    rc = L.require_device().dbl_dwconv2d(C.byref(xv), wd.data_ptr(), None if bd is None else bd.data_ptr(), C.byref(yv), k, s, d,
                                         act, stream())
    torch.cuda.synchronize()
    assert rc == 0, L.load().dbl_last_error().decode()
    assert_close_bf16(from_nhwc(ybuf, 0, c), ref, f"dwconv k{k} s{s} d{d}")


def test_pool_upsample_gate_copy():
    lib = L.require_device()
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.randn(2, 32, 16, 24, generator=g))
    b = bf16_round(torch.randn(2, 32, 16, 24, generator=g))
    xbuf, xv = to_nhwc(x, pad_c=16, c_off=8)
    bbuf, bv = to_nhwc(b)
    ybuf, yv = empty_nhwc(2, 8, 12, 32, pad_c=32, c_off=32)
    assert lib.dbl_avgpool2(C.byref(xv), C.byref(yv), stream()) == 0
    torch.cuda.synchronize()
    assert_close_bf16(from_nhwc(ybuf, 32, 32), F.avg_pool2d(x, 2), "avgpool2", abs_=1e-2)
    ybuf, yv = empty_nhwc(2, 32, 48, 32)
    assert lib.dbl_upsample2_nearest(C.byref(xv), C.byref(yv), stream()) == 0
    torch.cuda.synchronize()
    assert torch.equal(from_nhwc(ybuf, 0, 32), F.interpolate(x, scale_factor=2, mode="nearest"))
    ybuf, yv = empty_nhwc(2, 16, 24, 32)
    y2buf, y2v = empty_nhwc(2, 16, 24, 32, pad_c=8)
    assert lib.dbl_gated_add(C.byref(xv), C.byref(bv), C.c_float(0.5), C.byref(yv), C.byref(y2v), stream()) == 0
    torch.cuda.synchronize()
    assert_close_bf16(from_nhwc(ybuf, 0, 32), x + 0.5 * b, "gated_add", abs_=1e-2)
    assert torch.equal(from_nhwc(ybuf, 0, 32), from_nhwc(y2buf, 0, 32))
    ybuf, yv = empty_nhwc(2, 16, 24, 32)
    assert lib.dbl_copy(C.byref(xv), C.byref(yv), stream()) == 0
    torch.cuda.synchronize()
    assert torch.equal(from_nhwc(ybuf, 0, 32), x)


@pytest.mark.parametrize("c", [32, 64, 128])  # 8 channels per group -> 8-channel kernel; 16 / 32 per group -> 16-channel FFMA2 kernel
def test_dysample_vs_oracle(c):
    lib = L.require_device()
    g = torch.Generator().manual_seed(4)
    n, h, w = 2, 10, 14
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    off_raw = torch.randn(n, 32, h, w, generator=g) * 2.0  # large offsets: exercises the border clamp
    # oracle: feed the same raw offsets through DySample's own math by making its 1x1 conv the identity on a stub input
    sd = {"p.offset.weight": torch.zeros(32, c, 1, 1), "p.offset.bias": torch.zeros(32), "p.init_pos": O.dysample_init_pos()}
    orig = F.conv2d

    def fake_conv(inp, wt, bias=None, *a, **k):
        return off_raw if wt is sd["p.offset.weight"] else orig(inp, wt, bias, *a, **k)

    F.conv2d = fake_conv
    try:
        ref = O.dysample(sd, "p", x)
    finally:
        F.conv2d = orig
    xbuf, xv = to_nhwc(x)
    obuf, ov = to_nhwc(off_raw, dtype=torch.float32)
    ybuf, yv = empty_nhwc(n, 2 * h, 2 * w, c)
    assert lib.dbl_dysample(C.byref(xv), C.byref(ov), C.byref(yv), 4, stream()) == 0, lib.dbl_last_error().decode()
    torch.cuda.synchronize()
    assert_close_bf16(from_nhwc(ybuf, 0, c), ref, "dysample")


def test_lsk_squeeze_gate_vs_torch():
    lib = L.require_device()
    g = torch.Generator().manual_seed(5)
    n, half, h, w = 2, 64, 12, 20
    a = bf16_round(torch.randn(n, 2 * half, h, w, generator=g))
    wsq = torch.randn(2, 2, 7, 7, generator=g) * 0.2
    bsq = torch.randn(2, generator=g)
    agg = torch.cat([a.mean(1, keepdim=True), a.max(1, keepdim=True)[0]], 1)
    sig = F.conv2d(agg, wsq, bsq, 1, 3).sigmoid()
    ref = a[:, :half] * sig[:, 0:1] + a[:, half:] * sig[:, 1:2]
    abuf, av = to_nhwc(a)
    gbuf, gv = empty_nhwc(n, h, w, 2, torch.float32)
    ybuf, yv = empty_nhwc(n, h, w, half)
    assert lib.dbl_lsk_squeeze(C.byref(av), C.byref(gv), stream()) == 0
    wsd, bsd = wsq.reshape(-1).to(dev()), bsq.to(dev())
    assert lib.dbl_lsk_gate(C.byref(av), C.byref(gv), wsd.data_ptr(), bsd.data_ptr(), C.byref(yv), stream()) == 0
    torch.cuda.synchronize()
    torch.testing.assert_close(from_nhwc(gbuf, 0, 2), agg, rtol=1e-4, atol=1e-4)
    assert_close_bf16(from_nhwc(ybuf, 0, half), ref, "lsk gate")


@pytest.mark.parametrize("D,E,hw", [(64, 4, (20, 20)), (128, 8, (12, 10)), (256, 8, (8, 8))])
def test_hypergraph_vs_oracle(D, E, hw):
    lib = L.require_device()
    g = torch.Generator().manual_seed(6)
    B, (h, w) = 2, hw
    x = bf16_round(torch.randn(B, D, h, w, generator=g))
    p = "m.hgnn"
    sd = {
        p + ".edge_generator.prototype_base": torch.randn(E, D, generator=g) * 0.5,
        p + ".edge_generator.context_net.weight": torch.randn(E * D, 2 * D, generator=g) / (2 * D) ** 0.5,
        p + ".edge_generator.context_net.bias": torch.randn(E * D, generator=g) * 0.1,
        p + ".edge_generator.pre_head_proj.weight": torch.randn(D, D, generator=g) / D ** 0.5,
        p + ".edge_generator.pre_head_proj.bias": torch.randn(D, generator=g) * 0.1,
        p + ".edge_proj.0.weight": torch.randn(D, D, generator=g) / D ** 0.5, p + ".edge_proj.0.bias": torch.randn(D, generator=g) * 0.1,
        p + ".node_proj.0.weight": torch.randn(D, D, generator=g) / D ** 0.5, p + ".node_proj.0.bias": torch.randn(D, generator=g) * 0.1,
    }
    tok = O.hypergraph_conv(sd, p, x.flatten(2).transpose(1, 2), D // 16)
    ref = tok.transpose(1, 2).reshape(B, D, h, w)
    keep = {k: v.to(dev()).contiguous() for k, v in sd.items()}
    wt = L.HypergraphWeights(*[keep[p + s].data_ptr() for s in (
        ".edge_generator.prototype_base", ".edge_generator.context_net.weight", ".edge_generator.context_net.bias",
        ".edge_generator.pre_head_proj.weight", ".edge_generator.pre_head_proj.bias", ".edge_proj.0.weight", ".edge_proj.0.bias",
        ".node_proj.0.weight", ".node_proj.0.bias")], E, D // 16)
    xbuf, xv = to_nhwc(x, pad_c=D)  # tokens are a channel slice, as inside C3AH
    ybuf, yv = empty_nhwc(B, h, w, D)
    nbytes = lib.dbl_hypergraph_workspace_bytes(B, h * w, D, E)
    ws = torch.zeros(nbytes // 4 + 4, device=dev())
    rc = lib.dbl_hypergraph(C.byref(xv), C.byref(wt), C.byref(yv), ws.data_ptr(), C.c_size_t(nbytes), stream())
    torch.cuda.synchronize()
    assert rc == 0, lib.dbl_last_error().decode()
    assert_close_bf16(from_nhwc(ybuf, 0, D), ref, f"hypergraph D={D} E={E}")
    # in place (as the engine uses it)
    rc = lib.dbl_hypergraph(C.byref(xv), C.byref(wt), C.byref(xv), ws.data_ptr(), C.c_size_t(nbytes), stream())
    torch.cuda.synchronize()
    assert rc == 0
    assert_close_bf16(from_nhwc(xbuf, 0, D), ref, f"hypergraph in place D={D}")


def test_argument_errors_are_reported_not_thrown():
    lib = L.require_device()
    xbuf, xv = empty_nhwc(1, 8, 8, 16)
    ybuf, yv = empty_nhwc(1, 8, 8, 16)
    assert lib.dbl_conv2d(C.byref(xv), None, None, None, C.byref(yv), 3, 1, 1, 1, 0, 0, 0, stream()) == 1  # DBL_E_ARG
    assert b"null" in lib.dbl_last_error()
    bad = L.view(ybuf, 0, 16)
    bad.h = 9
    w = torch.zeros(16 * 9 * 16, dtype=torch.bfloat16, device=dev())
    assert lib.dbl_conv2d(C.byref(xv), w.data_ptr(), None, None, C.byref(bad), 3, 1, 1, 1, 0, 0, 0, stream()) == 1
    assert lib.dbl_conv2d(C.byref(xv), w.data_ptr(), None, None, C.byref(yv), 3, 1, 2, 2, 0, 0, L.IMPL_TCGEN05, stream()) == 2  # unsupported


# ---- warp-MMA kernels for the thin high-resolution layers (csrc/conv_mma.cu) -----------------------------------------
def _pack_mma(wt):
    """[Cout,Cin,3,3] fp32 -> device buffer in B-fragment order"""
    lib = L.require_device()
    cout, cin = wt.shape[:2]
    wd = pack_w(wt)
    nbytes = lib.dbl_mma_packed_bytes(cout, cin, 9)
    assert nbytes > 0
    out = torch.zeros(nbytes // 4, dtype=torch.int32, device=dev())
    assert lib.dbl_mma_pack_weights(wd.data_ptr(), cout, cin, 9, out.data_ptr(), stream()) == 0
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("c,n,h,w,xpad,ypad", [(16, 2, 40, 72, 0, 0), (16, 1, 33, 50, 16, 8), (32, 2, 48, 64, 0, 32), (32, 1, 17, 35, 0, 0),
                                               (64, 2, 24, 40, 64, 0), (64, 1, 9, 33, 0, 0), (16, 3, 320, 320, 0, 0)])
def test_bottleneck3x3_fused(c, n, h, w, xpad, ypad):
    g = torch.Generator().manual_seed(c + h)
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    w1 = bf16_round(torch.randn(c // 2, c, 3, 3, generator=g) / (9 * c) ** 0.5)
    w2 = bf16_round(torch.randn(c, c // 2, 3, 3, generator=g) / (4.5 * c) ** 0.5)
    b1, b2 = torch.randn(c // 2, generator=g) * 0.3, torch.randn(c, generator=g) * 0.3
    hid = bf16_round(F.silu(F.conv2d(x, w1, b1, 1, 1)))  # the hidden tensor is held as bf16 in shared memory
    ref = x + F.silu(F.conv2d(hid, w2, b2, 1, 1))
    xbuf, xv = to_nhwc(x, pad_c=xpad, c_off=xpad // 2 // 8 * 8)
    ybuf, yv = empty_nhwc(n, h, w, c, pad_c=ypad, c_off=ypad // 2 // 8 * 8)
    w1p, w2p, b1d, b2d = _pack_mma(w1), _pack_mma(w2), b1.to(dev()), b2.to(dev())
    lib = L.require_device()
    rc = lib.dbl_bottleneck3x3(C.byref(xv), w1p.data_ptr(), b1d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), 1, C.byref(yv), stream())
    torch.cuda.synchronize()
    assert rc == 0, L.load().dbl_last_error().decode()
    off = (yv.data - ybuf.data_ptr()) // 2
    assert_close_bf16(from_nhwc(ybuf, off, c), ref, f"bottleneck c={c} {h}x{w}", abs_=3e-2)
    if ypad:
        mask = torch.ones(ybuf.shape[-1], dtype=torch.bool)
        mask[off:off + c] = False
        assert bool((ybuf[..., mask.to(dev())].float() == -3.0).all()), "kernel wrote outside its channel slice"
    # shortcut = 0
    rc = lib.dbl_bottleneck3x3(C.byref(xv), w1p.data_ptr(), b1d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), 0, C.byref(yv), stream())
    torch.cuda.synchronize()
    assert rc == 0
    assert_close_bf16(from_nhwc(ybuf, off, c), ref - x, f"bottleneck (no shortcut) c={c}", abs_=3e-2)
    # unsupported widths and aliasing fail loudly
    assert lib.dbl_bottleneck3x3(C.byref(xv), w1p.data_ptr(), None, w2p.data_ptr(), None, 1, C.byref(xv), stream()) == 1


@pytest.mark.parametrize("c,n,h,w,ypad", [(16, 2, 40, 72, 0), (16, 1, 33, 52, 0), (16, 1, 33, 52, 16), (32, 2, 48, 64, 32), (32, 2, 48, 64, 0), (32, 1, 17, 34, 0), (16, 1, 7, 4, 0),
                                          (32, 1, 5, 2, 0), (16, 3, 320, 320, 0), (32, 6, 160, 160, 0)])
def test_bottleneck3x3_tc(c, n, h, w, ypad):
    """Bottleneck.forward on the tcgen05 pixel-group kernel (csrc/bottleneck_tc.cu) vs torch fp32 with the hidden tensor rounded to bf16"""
    g = torch.Generator().manual_seed(c + h)
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    w1 = bf16_round(torch.randn(c // 2, c, 3, 3, generator=g) / (9 * c) ** 0.5)
    w2 = bf16_round(torch.randn(c, c // 2, 3, 3, generator=g) / (4.5 * c) ** 0.5)
    b1, b2 = torch.randn(c // 2, generator=g) * 0.3, torch.randn(c, generator=g) * 0.3
    hid = bf16_round(F.silu(F.conv2d(x, w1, b1, 1, 1)))
    conv = F.silu(F.conv2d(hid, w2, b2, 1, 1))
    xbuf, xv = to_nhwc(x)
    ybuf, yv = empty_nhwc(n, h, w, c, pad_c=ypad, c_off=ypad // 2 // 8 * 8)
    lib = L.require_device()
    assert lib.dbl_bottleneck_tc_supported(c, w, xv.pitch) == 1
    packed = torch.zeros((lib.dbl_bottleneck_tc_packed_bytes(c) + 15) // 16 * 4, dtype=torch.int32, device=dev())
    b1d, b2d = b1.to(dev()), b2.to(dev())
    assert lib.dbl_bottleneck_tc_pack(pack_w(w1).data_ptr(), b1d.data_ptr(), pack_w(w2).data_ptr(), b2d.data_ptr(), c, packed.data_ptr(), stream()) == 0
    torch.cuda.synchronize()
    off = (yv.data - ybuf.data_ptr()) // 2
    if ypad:  # the output is stored through its pixel-group view as well: a channel slice is refused, not mis-written
        assert lib.dbl_bottleneck3x3_tc(C.byref(xv), packed.data_ptr(), 1, C.byref(yv), stream()) == 2  # DBL_E_UNSUPPORTED
        assert bool((ybuf.float() == -3.0).all())
        return
    for shortcut in (1, 0):
        rc = lib.dbl_bottleneck3x3_tc(C.byref(xv), packed.data_ptr(), shortcut, C.byref(yv), stream())
        torch.cuda.synchronize()
        assert rc == 0, L.load().dbl_last_error().decode()
        assert_close_bf16(from_nhwc(ybuf, off, c), (x + conv) if shortcut else conv, f"bottleneck_tc c={c} {h}x{w} shortcut={shortcut}", abs_=3e-2)
    # a channel slice as input has no pixel-group view
    xb2, xv2 = to_nhwc(x, pad_c=16, c_off=8)
    assert lib.dbl_bottleneck_tc_supported(c, w, xv2.pitch) == 0
    assert lib.dbl_bottleneck3x3_tc(C.byref(xv2), packed.data_ptr(), 1, C.byref(yv), stream()) == 2  # DBL_E_UNSUPPORTED


@pytest.mark.parametrize("cin,cout,s,n,h,w,epi", [(8, 16, 2, 2, 64, 96, L.EPI_NONE), (16, 32, 2, 2, 40, 72, L.EPI_NONE), (32, 64, 2, 1, 34, 38, L.EPI_NONE),
                                                  (16, 8, 1, 1, 33, 50, L.EPI_NONE), (8, 16, 1, 2, 24, 40, L.EPI_ADD), (32, 16, 1, 1, 40, 40, L.EPI_NONE),
                                                  (16, 32, 1, 1, 18, 70, L.EPI_MUL), (64, 32, 1, 1, 20, 36, L.EPI_NONE), (32, 64, 1, 2, 20, 20, L.EPI_ADD),
                                                  (16, 32, 2, 4, 160, 160, L.EPI_NONE)])
def test_conv3x3_mma(cin, cout, s, n, h, w, epi):
    g = torch.Generator().manual_seed(cin * 3 + cout)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.5
    ref = F.silu(F.conv2d(x, wt, b, s, 1))
    r = bf16_round(torch.randn(ref.shape, generator=g)) if epi else None
    if epi == L.EPI_ADD:
        ref = ref + r
    elif epi == L.EPI_MUL:
        ref = ref * r
    xbuf, xv = to_nhwc(x, pad_c=8, c_off=8)
    ybuf, yv = empty_nhwc(n, ref.shape[2], ref.shape[3], cout, pad_c=8)
    rv = None
    if epi:
        rbuf, rv = to_nhwc(r)
    wp, bd = _pack_mma(wt), b.to(dev())
    rc = L.require_device().dbl_conv3x3_mma(C.byref(xv), wp.data_ptr(), bd.data_ptr(), None if rv is None else C.byref(rv), C.byref(yv), s,
                                            L.ACT_SILU, epi, stream())
    torch.cuda.synchronize()
    assert rc == 0, L.load().dbl_last_error().decode()
    assert_close_bf16(from_nhwc(ybuf, 0, cout), ref, f"conv3x3_mma {cin}->{cout} s{s}")
    assert bool((ybuf[..., cout:].float() == -3.0).all())


@pytest.mark.parametrize("c0,c1", [(8, 16), (16, 32)])
@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("n,h,w", [(2, 64, 96), (1, 160, 224), (1, 34, 70)])
def test_stem_pair(c0, c1, u8, n, h, w):
    g = torch.Generator().manual_seed(c0 + h + int(u8))
    if u8:
        xi = torch.randint(0, 256, (n, 3, h, w), generator=g, dtype=torch.uint8)
        x = xi.float() / 255.0
    else:
        x = torch.rand(n, 3, h, w, generator=g)
    w0 = torch.randn(c0, 3, 3, 3, generator=g) / 27 ** 0.5
    b0 = torch.randn(c0, generator=g) * 0.3
    w1 = bf16_round(torch.randn(c1, c0, 3, 3, generator=g) / (9 * c0) ** 0.5)
    b1 = torch.randn(c1, generator=g) * 0.3
    l0 = bf16_round(F.silu(F.conv2d(x, w0, b0, 1, 1)))  # layer 0's output is held as bf16 in shared memory
    ref = F.silu(F.conv2d(l0, w1, b1, 2, 1))
    ybuf, yv = empty_nhwc(n, h // 2, w // 2, c1, pad_c=8)
    xd = (xi if u8 else x).to(dev())
    w0d, b0d, b1d, w1p = w0.to(dev()), b0.to(dev()), b1.to(dev()), _pack_mma(w1)
    rc = L.require_device().dbl_stem_pair(xd.data_ptr(), int(u8), n, h, w, w0d.data_ptr(), b0d.data_ptr(), c0, w1p.data_ptr(), b1d.data_ptr(),
                                          C.byref(yv), stream())
    torch.cuda.synchronize()
    assert rc == 0, L.load().dbl_last_error().decode()
    # layer 0 runs on bf16-rounded pixels / weights here (fp32 in the reference): one extra 2^-9 rounding per operand
    assert_close_bf16(from_nhwc(ybuf, 0, c1), ref, f"stem_pair {c0}->{c1} u8={u8}", rel=2.0 ** -6, abs_=4e-2)
    assert bool((ybuf[..., c1:].float() == -3.0).all())


@pytest.mark.parametrize("u8,n,h,w", [(False, 2, 64, 96), (True, 2, 64, 96), (False, 1, 34, 48), (True, 1, 34, 48), (True, 1, 2, 16), (False, 3, 128, 160),
                                      (True, 2, 640, 640)])
def test_stem_pair_tc(u8, n, h, w):
    """Layers 0+1 on the tcgen05 pixel-group kernel (csrc/stem_tc.cu) vs torch fp32 with layer 0's output rounded to bf16"""
    c0, c1 = 8, 16
    g = torch.Generator().manual_seed(h + int(u8))
    if u8:
        xi = torch.randint(0, 256, (n, 3, h, w), generator=g, dtype=torch.uint8)
        x = xi.float() / 255.0
    else:
        x = torch.rand(n, 3, h, w, generator=g)
    w0 = torch.randn(c0, 3, 3, 3, generator=g) / 27 ** 0.5
    b0 = torch.randn(c0, generator=g) * 0.3
    w1 = bf16_round(torch.randn(c1, c0, 3, 3, generator=g) / (9 * c0) ** 0.5)
    b1 = torch.randn(c1, generator=g) * 0.3
    l0 = bf16_round(F.silu(F.conv2d(x, w0, b0, 1, 1)))  # layer 0's output is held as bf16 in shared memory
    ref = F.silu(F.conv2d(l0, w1, b1, 2, 1))
    lib = L.require_device()
    ybuf, yv = empty_nhwc(n, h // 2, w // 2, c1)
    assert lib.dbl_stem_tc_supported(c0, c1, h, w, yv.pitch) == 1
    xd = (xi if u8 else x).to(dev())
    w0d, b0d, b1d, w1d = w0.to(dev()), b0.to(dev()), b1.to(dev()), pack_w(w1)
    packed = torch.zeros((lib.dbl_stem_tc_packed_bytes() + 15) // 16 * 4, dtype=torch.int32, device=dev())
    assert lib.dbl_stem_tc_pack(w0d.data_ptr(), b0d.data_ptr(), w1d.data_ptr(), b1d.data_ptr(), int(u8), packed.data_ptr(), stream()) == 0
    rc = lib.dbl_stem_pair_tc(xd.data_ptr(), int(u8), n, h, w, packed.data_ptr(), C.byref(yv), stream())
    torch.cuda.synchronize()
    assert rc == 0, L.load().dbl_last_error().decode()
    # layer 0 runs on bf16-rounded pixels / weights here (fp32 in the reference): one extra 2^-9 rounding per operand
    assert_close_bf16(from_nhwc(ybuf, 0, c1), ref, f"stem_pair_tc u8={u8} {h}x{w}", rel=2.0 ** -6, abs_=4e-2)
    # a channel-slice output has no pixel-group view: refused, not mis-written
    yb2, yv2 = empty_nhwc(n, h // 2, w // 2, c1, pad_c=8)
    assert lib.dbl_stem_tc_supported(c0, c1, h, w, yv2.pitch) == 0
    assert lib.dbl_stem_pair_tc(xd.data_ptr(), int(u8), n, h, w, packed.data_ptr(), C.byref(yv2), stream()) == 2  # DBL_E_UNSUPPORTED


# ---- multi-accumulator tcgen05 3x3 kernel (csrc/conv_tc4.cu): large reduction depth --------------------------------
TC4_CASES = [
    (2, 40, 40, 384, 64, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 80, 80, 256, 32, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (2, 40, 40, 192, 64, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 64),
    (2, 20, 20, 256, 64, 3, 1, L.ACT_NONE, L.EPI_NONE, torch.float32, 0, 0),
    (1, 40, 40, 128, 128, 3, 1, L.ACT_SILU, L.EPI_ADD, torch.bfloat16, 128, 0),
    (1, 33, 50, 128, 48, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (1, 20, 20, 160, 80, 3, 1, L.ACT_GELU, L.EPI_MUL, torch.bfloat16, 0, 0),   # 32-channel chunks
    (1, 24, 24, 144, 32, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),   # 16-channel chunks
    (1, 20, 20, 256, 256, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
    (24, 80, 80, 128, 32, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),  # every CTA loops over several super-tiles
    (3, 40, 40, 64, 64, 3, 1, L.ACT_SILU, L.EPI_NONE, torch.bfloat16, 0, 0),
]


@pytest.mark.parametrize("case", TC4_CASES)
def test_conv3x3_multi_accumulator_kernel(case):
    _run_conv_case(case, L.IMPL_TCGEN05_MULTI, seed=11)


# ---- fused depthwise-separable convolution (csrc/dsconv_mma.cu) -----------------------------------------------------
@pytest.mark.parametrize("k,s,cin,cout,dw_act,epi,n,h,w", [
    (3, 1, 64, 64, 0, L.EPI_NONE, 2, 40, 40), (7, 1, 64, 64, 0, L.EPI_ADD, 2, 40, 40), (3, 1, 32, 32, 0, L.EPI_NONE, 1, 24, 36),
    (7, 1, 32, 32, 0, L.EPI_ADD, 1, 40, 40), (3, 1, 128, 128, 0, L.EPI_NONE, 2, 20, 20), (7, 1, 128, 128, 0, L.EPI_ADD, 1, 20, 20),
    (3, 2, 128, 128, 0, L.EPI_NONE, 1, 80, 80), (3, 2, 128, 256, 0, L.EPI_NONE, 2, 40, 40), (3, 1, 64, 80, 1, L.EPI_NONE, 1, 80, 80),
    (3, 1, 80, 80, 1, L.EPI_NONE, 2, 40, 40), (3, 1, 128, 80, 1, L.EPI_NONE, 1, 40, 40), (3, 1, 256, 80, 1, L.EPI_NONE, 2, 20, 20),
    (3, 1, 128, 64, 1, L.EPI_NONE, 1, 13, 19), (3, 1, 256, 64, 1, L.EPI_NONE, 1, 10, 10), (7, 1, 64, 64, 0, L.EPI_ADD, 1, 33, 50)])
def test_dsconv_fused(k, s, cin, cout, dw_act, epi, n, h, w):
    lib = L.require_device()
    assert lib.dbl_dsconv_supported(k, s, cin, cout) == 1
    g = torch.Generator().manual_seed(k * 100 + cin + cout)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wdw = torch.randn(cin, 1, k, k, generator=g) / k
    bdw = torch.randn(cin, generator=g) * 0.3 if dw_act else None
    wpw = bf16_round(torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5)
    bpw = torch.randn(cout, generator=g) * 0.3
    d = F.conv2d(x, wdw, bdw, s, (k - 1) // 2, 1, cin)
    if dw_act:
        d = F.silu(d)
    d = bf16_round(d)  # the depthwise result is held as bf16 in shared memory
    ref = F.silu(F.conv2d(d, wpw, bpw))
    r = bf16_round(torch.randn(ref.shape, generator=g)) if epi else None
    if epi == L.EPI_ADD:
        ref = ref + r
    xbuf, xv = to_nhwc(x, pad_c=8, c_off=0)
    ybuf, yv = empty_nhwc(n, ref.shape[2], ref.shape[3], cout, pad_c=16, c_off=8)
    rv = None
    if epi:
        rbuf, rv = to_nhwc(r)
    wd = wdw.reshape(cin, k * k).t().contiguous().to(dev())
    bd = None if bdw is None else bdw.to(dev())
    cout_, cin_ = wpw.shape[:2]
    wdev = wpw.permute(0, 2, 3, 1).contiguous().to(dev(), torch.bfloat16)
    nbytes = lib.dbl_mma_packed_bytes(cout_, cin_, 1)
    wp = torch.zeros(nbytes // 4, dtype=torch.int32, device=dev())
    assert lib.dbl_mma_pack_weights(wdev.data_ptr(), cout_, cin_, 1, wp.data_ptr(), stream()) == 0
    bp = bpw.to(dev())
    rc = lib.dbl_dsconv(C.byref(xv), wd.data_ptr(), None if bd is None else bd.data_ptr(), dw_act, wp.data_ptr(), bp.data_ptr(),
                        None if rv is None else C.byref(rv), C.byref(yv), k, s, epi, stream())
    torch.cuda.synchronize()
    assert rc == 0, L.load().dbl_last_error().decode()
    assert_close_bf16(from_nhwc(ybuf, 8, cout), ref, f"dsconv k{k} s{s} {cin}->{cout}", abs_=3e-2)
    assert bool((ybuf[..., :8].float() == -3.0).all()) and bool((ybuf[..., 8 + cout:].float() == -3.0).all())
    assert lib.dbl_dsconv_supported(5, 1, 48, 24) == 0


# ---- depthwise-separable convolution on the Blackwell data path (csrc/dsconv_tc.cu: TMA + tcgen05 + TMEM, persistent) -----
@pytest.mark.parametrize("k,cin,cout,dw_act,epi,n,h,w", [
    (3, 64, 64, 0, L.EPI_NONE, 2, 40, 40), (7, 64, 64, 0, L.EPI_ADD, 2, 40, 40), (7, 64, 64, 0, L.EPI_ADD, 12, 40, 40),
    (3, 32, 32, 0, L.EPI_NONE, 1, 24, 36), (7, 32, 32, 0, L.EPI_ADD, 3, 40, 40), (3, 128, 128, 0, L.EPI_NONE, 2, 20, 20),
    (7, 128, 128, 0, L.EPI_ADD, 5, 20, 20), (3, 64, 80, 1, L.EPI_NONE, 1, 80, 80), (3, 80, 80, 1, L.EPI_NONE, 2, 40, 40),
    (3, 128, 80, 1, L.EPI_NONE, 1, 40, 40), (3, 128, 64, 1, L.EPI_NONE, 1, 13, 19), (7, 64, 64, 0, L.EPI_ADD, 1, 33, 50),
    (3, 64, 64, 0, L.EPI_MUL, 1, 7, 9), (7, 64, 64, 0, L.EPI_NONE, 40, 20, 20), (3, 256, 80, 1, L.EPI_NONE, 2, 20, 20),
    (3, 256, 64, 1, L.EPI_NONE, 1, 10, 10)])
def test_dsconv_tc(k, cin, cout, dw_act, epi, n, h, w):
    _run_dsconv_tc(k, 1, cin, cout, dw_act, epi, n, h, w)


@pytest.mark.parametrize("n,h,w", [(2, 80, 80), (1, 33, 50), (3, 40, 24), (1, 6, 10)])
def test_dsconv_tc_stride2(n, h, w):
    """DSConv [512, 3, 2] of the backbone (layer 8 at n scale: 128 -> 128): the strided instantiation, even and odd map sizes"""
    _run_dsconv_tc(3, 2, 128, 128, 0, L.EPI_NONE, n, h, w)


def _run_dsconv_tc(k, s, cin, cout, dw_act, epi, n, h, w):
    """every (k, cin, cout) instantiation; tiles that hang over the image on every side (TMA zero fill / store clipping), more tiles
    than SMs (persistent loop, both pipeline parities), channel-slice views with a foreign pitch, and the engine's in-place residual
    (r aliases y)."""
    lib = L.require_device()
    assert lib.dbl_dsconv_tc_supported(k, s, cin, cout) == 1 and lib.dbl_dsconv_tc_supported(k, 3, cin, cout) == 0
    g = torch.Generator().manual_seed(k * 100 + cin + cout + n)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wdw = torch.randn(cin, 1, k, k, generator=g) / k
    bdw = torch.randn(cin, generator=g) * 0.3 if dw_act else None
    wpw = bf16_round(torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5)
    bpw = torch.randn(cout, generator=g) * 0.3
    d = F.conv2d(x, wdw, bdw, s, (k - 1) // 2, 1, cin)
    if dw_act:
        d = F.silu(d)
    d = bf16_round(d)  # the depthwise result is held as bf16 in shared memory (the A operand of the pointwise GEMM)
    ref = F.silu(F.conv2d(d, wpw, bpw))
    r = bf16_round(torch.randn(ref.shape, generator=g)) if epi else None
    if epi == L.EPI_ADD:
        ref = ref + r
    elif epi == L.EPI_MUL:
        ref = ref * r
    xbuf, xv = to_nhwc(x, pad_c=8, c_off=0)
    ybuf, yv = empty_nhwc(n, ref.shape[2], ref.shape[3], cout, pad_c=16, c_off=8)
    rv = None
    if epi:  # in place, like DSBottleneck's `x = x + cv2(cv1(x))` in the engine: the residual IS the output view
        ybuf[..., 8:8 + cout] = r.permute(0, 2, 3, 1).to(dev(), torch.bfloat16)
        rv = yv
    wd = wdw.reshape(cin, k * k).t().contiguous().to(dev())
    bd = None if bdw is None else bdw.to(dev())
    wdev = wpw.reshape(cout, cin).contiguous().to(dev(), torch.bfloat16)
    bp = bpw.to(dev())
    rc = lib.dbl_dsconv_tc(C.byref(xv), wd.data_ptr(), None if bd is None else bd.data_ptr(), dw_act, wdev.data_ptr(), bp.data_ptr(),
                           None if rv is None else C.byref(rv), C.byref(yv), k, epi, stream())
    torch.cuda.synchronize()
    assert rc == 0, L.load().dbl_last_error().decode()
    assert_close_bf16(from_nhwc(ybuf, 8, cout), ref, f"dsconv_tc k{k} {cin}->{cout}", abs_=3e-2)
    assert bool((ybuf[..., :8].float() == -3.0).all()) and bool((ybuf[..., 8 + cout:].float() == -3.0).all())  # neighbours of the view untouched
    assert bool((xbuf[..., cin:].float() == 7.0).all())
    # run-to-run determinism of the pipelined kernel
    if not epi:
        y1 = ybuf.clone()
        assert lib.dbl_dsconv_tc(C.byref(xv), wd.data_ptr(), None if bd is None else bd.data_ptr(), dw_act, wdev.data_ptr(), bp.data_ptr(), None,
                                 C.byref(yv), k, epi, stream()) == 0
        torch.cuda.synchronize()
        assert torch.equal(y1, ybuf)


# ---- area attention of the stock yolov13.yaml (csrc/attention.cu) ---------------------------------------------------------
@pytest.mark.parametrize("n,h,w,heads,area", [(2, 40, 40, 2, 4), (3, 20, 20, 4, 1), (1, 16, 20, 2, 4), (1, 8, 10, 1, 1), (2, 12, 25, 2, 4)])
def test_area_attention(n, h, w, heads, area):
    """AAttn.forward between the qkv conv and pe / proj (reference block.py:1245-1262), restated with torch fp32 on bf16-rounded inputs"""
    lib = L.require_device()
    hd = 32
    c = heads * hd
    g = torch.Generator().manual_seed(n * 1000 + h * w + heads)
    qkv = bf16_round(torch.randn(n, 3 * c, h, w, generator=g) * 1.5)
    B, N = n, h * w
    t = qkv.flatten(2).transpose(1, 2)
    if area > 1:
        t = t.reshape(B * area, N // area, 3 * c)
    Bq, Nq = t.shape[:2]
    q, k, v = t.view(Bq, Nq, heads, hd * 3).permute(0, 2, 3, 1).split([hd, hd, hd], dim=2)
    attn = ((q.transpose(-2, -1) @ k) * (hd ** -0.5)).softmax(dim=-1)
    o = (v @ attn.transpose(-2, -1)).permute(0, 3, 1, 2)
    vv = v.permute(0, 3, 1, 2)
    if area > 1:
        o, vv = o.reshape(B, N, c), vv.reshape(B, N, c)
    o = o.reshape(B, h, w, c).permute(0, 3, 1, 2)
    vv = vv.reshape(B, h, w, c).permute(0, 3, 1, 2)
    qbuf, qv = to_nhwc(qkv, pad_c=8)
    obuf, ov = empty_nhwc(n, h, w, c, pad_c=16, c_off=8)
    vbuf, vview = empty_nhwc(n, h, w, c)
    rc = lib.dbl_area_attention(C.byref(qv), C.byref(ov), C.byref(vview), heads, area, stream())
    torch.cuda.synchronize()
    assert rc == 0, L.load().dbl_last_error().decode()
    assert_close_bf16(from_nhwc(obuf, 8, c), o, f"area attention {h}x{w} heads {heads} area {area}", rel=2.0 ** -6, abs_=2e-2)
    assert torch.equal(from_nhwc(vbuf, 0, c), vv.contiguous())  # v is a copy
    assert bool((obuf[..., :8].float() == -3.0).all()) and bool((obuf[..., 8 + c:].float() == -3.0).all())
    assert lib.dbl_area_attention(C.byref(qv), C.byref(ov), C.byref(vview), heads + 1, area, stream()) != 0  # head dim != 32
