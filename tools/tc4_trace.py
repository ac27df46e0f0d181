"""Developer aid: time line of conv3x3_tc4_kernel CTAs (needs a library built with `make TRACE=1`)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

SHAPES = {"l6cv1_64to32@80": (64, 80, 80, 64, 32, 0), "l6cv2_32to64@80": (64, 80, 80, 32, 64, 1), "head_64to64@80": (64, 80, 80, 64, 64, 0)}


def main():
    lib = L.require_device()
    lib.dbl_debug_trace4_read.argtypes = [C.c_void_p, C.c_size_t]
    dev = torch.device("cuda:0")
    st = L.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, (n, h, w, cin, cout, epi) in SHAPES.items():
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        wt = (torch.randn(cout, 3, 3, cin, device=dev) / (9 * cin) ** 0.5).to(torch.bfloat16)
        b = torch.randn(cout, device=dev)
        y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
        r = torch.randn(n, h, w, cout, device=dev).to(torch.bfloat16)
        xv, yv, rv = L.view(x), L.view(y), L.view(r)
        args = (C.byref(xv), wt.data_ptr(), b.data_ptr(), C.byref(rv) if epi else None, C.byref(yv), 3, 1, 1, 1, L.ACT_SILU, epi, L.IMPL_TCGEN05_MULTI, st)
        for _ in range(2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            assert lib.dbl_conv2d(*args) == 0, lib.dbl_last_error().decode()
            e1.record()
            torch.cuda.synchronize()
        buf = np.zeros(160 * 64, dtype=np.uint64)
        assert lib.dbl_debug_trace4_read(buf.ctypes.data, buf.nbytes) == 0
        t = buf.reshape(160, 64)[:148].astype(np.int64)
        t0 = t[:, 0].min()
        print(f"{name}: event time {e0.elapsed_time(e1) * 1e3:.1f} us; CTA 0 / CTA 77 time line (us since first entry)")
        for cta in (0, 77):
            rel = (t[cta] - t0) / 1e3
            print(f"  cta {cta}: entry {rel[0]:.2f} pdl {rel[1]:.2f} exit {rel[63]:.2f}")
            print("    halo landed (MMA unit k): " + " ".join(f"{v:.2f}" for v in rel[2:32] if v > 0))
            print("    quad0 tile ready/done:    " + " ".join(f"{rel[32 + 2 * j]:.2f}/{rel[33 + 2 * j]:.2f}" for j in range(15) if rel[32 + 2 * j] > 0))
        print(f"  all CTAs: exit median {np.median((t[:, 63] - t0) / 1e3):.2f} max {((t[:, 63] - t0) / 1e3).max():.2f}")


if __name__ == "__main__":
    main()
