"""Micro-benchmark of dbl_conv2d on the layer shapes of YOLO-DBL-n at batch 64 (CUDA events, L2 flushed between reps).

    python tools/conv_bench.py [--only NAME] [--reps 5] [--impl 2]
"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

SHAPES = {  # name: (n, h, w, cin, cout, k, s, epi, ydtype)
    "l2cv1_16to8_320": (64, 320, 320, 16, 8, 3, 1, 0, torch.bfloat16),
    "l3_16to32_s2_320": (64, 320, 320, 16, 32, 3, 2, 0, torch.bfloat16),
    "l4cv1_32to16_160": (64, 160, 160, 32, 16, 3, 1, 0, torch.bfloat16),
    "l4cv2_16to32_160": (64, 160, 160, 16, 32, 3, 1, 1, torch.bfloat16),
    "l5_32to64_s2_160": (64, 160, 160, 32, 64, 3, 2, 0, torch.bfloat16),
    "l6cv1_64to32_80": (64, 80, 80, 64, 32, 3, 1, 0, torch.bfloat16),
    "l6cv2_32to64_80": (64, 80, 80, 32, 64, 3, 1, 1, torch.bfloat16),
    "l7_64to128_1x1_80": (64, 80, 80, 64, 128, 1, 1, 0, torch.bfloat16),
    "l24cv1_256to32_80": (64, 80, 80, 256, 32, 3, 1, 0, torch.bfloat16),
    "l20cv1_384to64_40": (64, 40, 40, 384, 64, 3, 1, 0, torch.bfloat16),
    "l20cv2_64to128_40": (64, 40, 40, 64, 128, 3, 1, 0, torch.bfloat16),
    "pw_64to64_1x1_40": (64, 40, 40, 64, 64, 1, 1, 1, torch.bfloat16),
    "c2f_384to128_1x1_40": (64, 40, 40, 384, 128, 1, 1, 0, torch.bfloat16),
    "l31_128to128_s2_40": (64, 40, 40, 128, 128, 3, 2, 0, torch.bfloat16),
    "p5_384to256_1x1_20": (64, 20, 20, 384, 256, 1, 1, 0, torch.bfloat16),
    "head_64to80_f32_80": (64, 80, 80, 80, 80, 1, 1, 0, torch.float32),
    "lsk_128to256_mul_20": (64, 20, 20, 128, 256, 1, 1, 2, torch.bfloat16),
    "lsk_128to256_20": (64, 20, 20, 128, 256, 1, 1, 0, torch.bfloat16),
    "lsk_256to128_20": (64, 20, 20, 256, 128, 1, 1, 0, torch.bfloat16),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--impl", type=int, default=L.IMPL_TCGEN05)
    ap.add_argument("--n", type=int, default=0, help="override the batch size of every shape")
    ap.add_argument("--act", type=int, default=L.ACT_SILU, help="0 none, 1 SiLU, 2 GELU")
    ap.add_argument("--warm", action="store_true", help="no L2 flush between repetitions")
    a = ap.parse_args()
    lib = L.require_device()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = L.current_stream()
    for name, (n, h, w, cin, cout, k, s, epi, ydt) in SHAPES.items():
        if a.only and a.only not in name:
            continue
        n = a.n or n
        ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        wt = (torch.randn(cout, k, k, cin, device=dev) / (cin * k * k) ** 0.5).to(torch.bfloat16)
        b = torch.randn(cout, device=dev)
        y = torch.empty(n, ho, wo, cout, dtype=ydt, device=dev)
        r = torch.randn(n, ho, wo, cout, device=dev).to(torch.bfloat16) if epi else None
        xv, yv = L.view(x), L.view(y)
        rv = L.view(r) if epi else None
        args = (C.byref(xv), wt.data_ptr(), b.data_ptr(), C.byref(rv) if epi else None, C.byref(yv), k, s, k // 2, 1, a.act, epi, a.impl, st)
        assert lib.dbl_conv2d(*args) == 0, lib.dbl_last_error().decode()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.reps):
            if not a.warm:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.dbl_conv2d(*args)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / a.reps
        nbytes = x.numel() * 2 + y.numel() * y.element_size() + wt.numel() * 2 + (r.numel() * 2 if epi else 0)
        flops = 2.0 * n * ho * wo * cout * cin * k * k
        print(f"{name:24s} {ms:8.4f} ms  {nbytes / ms / 1e6:8.0f} GB/s ({nbytes / ms / 1e6 / 6455.3 * 100:5.1f}% HBM)  {flops / ms / 1e9:8.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
