"""Developer aid: phase time line of conv_tc_kernel CTAs (needs a library built with `make TRACE=1`).

    python tools/tc_trace.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

SHAPES = {  # name: (n, h, w, cin, cout)
    "lsk_256to128@20": (64, 20, 20, 256, 128),
    "pw_64to64@40": (64, 40, 40, 64, 64),
    "pw_128to128@20": (64, 20, 20, 128, 128),
    "l7_64to128@80": (64, 80, 80, 64, 128),
    "c2f_384to128@40": (64, 40, 40, 384, 128),
}
NAMES = ["entry", "prologue done", "pdl_wait done", "first TMA issued", "first stage full", "last stage full (tile 0)",
         "tile 0 accumulated", "tile 0 stored", "last tile stored", "exit"]


def main():
    alt = os.path.join(os.path.dirname(L.LIB_PATH), "..", "lib_trace", "libdbl_sm100.so")  # a -DDBL_TRACE build kept next to the shipped one
    if os.path.exists(alt):
        L.LIB_PATH = alt
    lib = L.require_device()
    lib.dbl_debug_trace_read.argtypes = [C.c_void_p, C.c_size_t]
    dev = torch.device("cuda:0")
    st = L.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, (n, h, w, cin, cout) in SHAPES.items():
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        wt = (torch.randn(cout, 1, 1, cin, device=dev) / cin ** 0.5).to(torch.bfloat16)
        b = torch.randn(cout, device=dev)
        y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
        xv, yv = L.view(x), L.view(y)
        args = (C.byref(xv), wt.data_ptr(), b.data_ptr(), None, C.byref(yv), 1, 1, 0, 1, L.ACT_SILU, 0, L.IMPL_TCGEN05, st)
        for cold in (1, 0):
            for _ in range(2):
                if cold:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                assert lib.dbl_conv2d(*args) == 0, lib.dbl_last_error().decode()
                e1.record()
                torch.cuda.synchronize()
            buf = np.zeros(512 * 16, dtype=np.uint64)
            assert lib.dbl_debug_trace_read(buf.ctypes.data, buf.nbytes) == 0
            t = buf.reshape(512, 16)[:, :10].astype(np.int64)
            tiles = (n * h * w + 127) // 128
            ctas = min(tiles, 296)
            t = t[:ctas]
            t0 = t[:, 0].min()
            rel = (t - t0) / 1e3
            print(f"{name} ({'cold' if cold else 'L2-warm'}): event time {e0.elapsed_time(e1) * 1e3:.1f} us, {tiles} tiles, {ctas} CTAs; us since first CTA entry: median / max over CTAs")
            for i, nm in enumerate(NAMES):
                print(f"    {nm:26s} {np.median(rel[:, i]):7.2f} {rel[:, i].max():7.2f}")


if __name__ == "__main__":
    main()
