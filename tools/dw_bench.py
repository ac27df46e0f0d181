"""Micro-benchmark of dbl_dwconv2d on the depthwise shapes of YOLO-DBL-n at batch 64 (CUDA events).

    python tools/dw_bench.py            # heuristic tile
    DBL_DW_TILE=8,8,4 python tools/dw_bench.py   # forced (G, tw, ty)
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

SHAPES = {  # name: (n, h, w, c, pitch, k, s, d)
    "l9_k3_64@40": (64, 40, 40, 64, 128, 3, 1, 1),
    "l9_k7_64@40": (64, 40, 40, 64, 128, 7, 1, 1),
    "l8_k3s2_128@80": (64, 80, 80, 128, 128, 3, 2, 1),
    "lsk_k5_256@20": (64, 20, 20, 256, 256, 5, 1, 1),
    "lsk_k7d3_256@20": (64, 20, 20, 256, 256, 7, 1, 3),
    "det_k3_64@80": (64, 80, 80, 64, 64, 3, 1, 1),
    "det_k3_80@80": (64, 80, 80, 80, 80, 3, 1, 1),
    "l12_k7_32@40": (64, 40, 40, 32, 64, 7, 1, 1),
    "l33_k7_128@20": (64, 20, 20, 128, 256, 7, 1, 1),
}


def main():
    lib = L.require_device()
    dev = torch.device("cuda:0")
    st = L.current_stream()
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    out = []
    for name, (n, h, w, c, pitch, k, s, d) in SHAPES.items():
        if only and only not in name:
            continue
        pad = d * (k - 1) // 2
        ho, wo = (h + 2 * pad - d * (k - 1) - 1) // s + 1, (w + 2 * pad - d * (k - 1) - 1) // s + 1
        x = torch.randn(n, h, w, pitch, device=dev).to(torch.bfloat16)
        y = torch.empty(n, ho, wo, pitch, dtype=torch.bfloat16, device=dev)
        wt = torch.randn(k * k, c, device=dev)
        xv, yv = L.view(x, 0, c), L.view(y, 0, c)
        args = (C.byref(xv), wt.data_ptr(), None, C.byref(yv), k, s, d, 0, st)
        assert lib.dbl_dwconv2d(*args) == 0, lib.dbl_last_error().decode()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            lib.dbl_dwconv2d(*args)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        out.append(f"{name}={us:.1f}")
    print(os.environ.get("DBL_DW_TILE", "auto"), " ".join(out))


if __name__ == "__main__":
    main()
