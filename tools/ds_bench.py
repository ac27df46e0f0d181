"""Micro-benchmark of the two fused depthwise-separable kernels -- dbl_dsconv (warp MMA, csrc/dsconv_mma.cu) and dbl_dsconv_tc (TMA +
tcgen05 + TMEM, persistent, csrc/dsconv_tc.cu) -- on the DSBottleneck / Detect class-branch shapes of YOLO-DBL-n at batch 64.

    python tools/ds_bench.py [name-substring]        DBL_DS_TILE=tw,th forces the tile of the warp-MMA kernel
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

SHAPES = {  # name: (n, h, w, cin, cout, k, epi, dw_act[, stride])
    "l8_k3s2_128@80": (64, 80, 80, 128, 128, 3, 0, 0, 2),
    "l9_k3_64@40": (64, 40, 40, 64, 64, 3, 0, 0),
    "l9_k7_64@40": (64, 40, 40, 64, 64, 7, 1, 0),
    "l12_k3_32@40": (64, 40, 40, 32, 32, 3, 0, 0),
    "l12_k7_32@40": (64, 40, 40, 32, 32, 7, 1, 0),
    "l33_k3_128@20": (64, 20, 20, 128, 128, 3, 0, 0),
    "l33_k7_128@20": (64, 20, 20, 128, 128, 7, 1, 0),
    "det_p3_64to80@80": (64, 80, 80, 64, 80, 3, 0, 1),
    "det_p3_80to80@80": (64, 80, 80, 80, 80, 3, 0, 1),
    "det_p4_128to80@40": (64, 40, 40, 128, 80, 3, 0, 1),
    "det_p4_80to80@40": (64, 40, 40, 80, 80, 3, 0, 1),
    "det_p5_256to80@20": (64, 20, 20, 256, 80, 3, 0, 1),
    "det_p5_80to80@20": (64, 20, 20, 80, 80, 3, 0, 1),
}


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    lib = L.require_device()
    dev = torch.device("cuda:0")
    st = L.current_stream()
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, shp in SHAPES.items():
        n, h, w, cin, cout, k, epi, dw_act = shp[:8]
        st_ = shp[8] if len(shp) > 8 else 1
        if only and only not in name:
            continue
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        y = torch.empty(n, h // st_, w // st_, cout, dtype=torch.bfloat16, device=dev)
        r = torch.randn(n, h // st_, w // st_, cout, device=dev).to(torch.bfloat16)
        wd = torch.randn(k * k, cin, device=dev) / k
        bd = torch.randn(cin, device=dev) * 0.1
        wpw = (torch.randn(cout, 1, 1, cin, device=dev) / cin ** 0.5).to(torch.bfloat16)
        bp = torch.randn(cout, device=dev)
        xv, yv, rv = L.view(x), L.view(y), L.view(r)
        res = {}
        if lib.dbl_dsconv_supported(k, st_, cin, cout):
            wp = torch.zeros(lib.dbl_mma_packed_bytes(cout, cin, 1) // 4, dtype=torch.int32, device=dev)
            assert lib.dbl_mma_pack_weights(wpw.data_ptr(), cout, cin, 1, wp.data_ptr(), st) == 0
            args = (C.byref(xv), wd.data_ptr(), bd.data_ptr() if dw_act else None, dw_act, wp.data_ptr(), bp.data_ptr(), C.byref(rv) if epi else None, C.byref(yv), k, st_, epi, st)
            assert lib.dbl_dsconv(*args) == 0, lib.dbl_last_error().decode()
            res["mma"] = timeit(lambda: lib.dbl_dsconv(*args))
            y_mma = y.clone()
        if lib.dbl_dsconv_tc_supported(k, st_, cin, cout):
            args2 = (C.byref(xv), wd.data_ptr(), bd.data_ptr() if dw_act else None, dw_act, wpw.data_ptr(), bp.data_ptr(), C.byref(rv) if epi else None, C.byref(yv), k, epi, st)
            assert lib.dbl_dsconv_tc(*args2) == 0, lib.dbl_last_error().decode()
            res["tc"] = timeit(lambda: lib.dbl_dsconv_tc(*args2))
            if "mma" in res:
                res["maxdiff"] = float((y.float() - y_mma.float()).abs().max())
        nbytes = 2 * n * (h * w * cin + (h // st_) * (w // st_) * cout * (2 if epi else 1))
        flo = {kk: f"{v:.1f}us ({nbytes / v / 1e3:.0f} GB/s)" if kk != "maxdiff" else f"{v:.3g}" for kk, v in res.items()}
        print(f"{name:20s} {flo}", flush=True)


if __name__ == "__main__":
    main()
