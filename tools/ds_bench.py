"""Micro-benchmark of dbl_dsconv (fused depthwise + pointwise) on the DSBottleneck shapes of YOLO-DBL-n at batch 64.

    python tools/ds_bench.py [name-substring]        DBL_DS_TILE=tw,th forces the tile
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

SHAPES = {  # name: (n, h, w, cin, cout, k, epi)
    "l9_k3_64@40": (64, 40, 40, 64, 64, 3, 0),
    "l9_k7_64@40": (64, 40, 40, 64, 64, 7, 1),
    "l12_k7_32@40": (64, 40, 40, 32, 32, 7, 1),
    "l33_k3_128@20": (64, 20, 20, 128, 128, 3, 0),
    "l33_k7_128@20": (64, 20, 20, 128, 128, 7, 1),
}


def main():
    lib = L.require_device()
    dev = torch.device("cuda:0")
    st = L.current_stream()
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    out = []
    for name, (n, h, w, cin, cout, k, epi) in SHAPES.items():
        if only and only not in name:
            continue
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
        r = torch.randn(n, h, w, cout, device=dev).to(torch.bfloat16)
        wd = torch.randn(k * k, cin, device=dev) / k
        wpw = (torch.randn(cout, 1, 1, cin, device=dev) / cin ** 0.5).to(torch.bfloat16)
        wp = torch.zeros(lib.dbl_mma_packed_bytes(cout, cin, 1) // 4, dtype=torch.int32, device=dev)
        assert lib.dbl_mma_pack_weights(wpw.data_ptr(), cout, cin, 1, wp.data_ptr(), st) == 0
        bp = torch.randn(cout, device=dev)
        xv, yv, rv = L.view(x), L.view(y), L.view(r)
        args = (C.byref(xv), wd.data_ptr(), None, 0, wp.data_ptr(), bp.data_ptr(), C.byref(rv) if epi else None, C.byref(yv), k, 1, epi, st)
        assert lib.dbl_dsconv(*args) == 0, lib.dbl_last_error().decode()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            lib.dbl_dsconv(*args)
        e1.record()
        torch.cuda.synchronize()
        out.append(f"{name}={e0.elapsed_time(e1) / reps * 1e3:.1f}")
    print(os.environ.get("DBL_DS_TILE", "auto"), " ".join(out), flush=True)


if __name__ == "__main__":
    main()
