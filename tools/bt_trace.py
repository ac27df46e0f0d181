"""Developer aid: time line of the roles of bottleneck_tc_kernel's CTA 0 (producer / MMA / epilogue), first tiles.

    python tools/bt_trace.py [c] [hw]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

ROLES = ["producer", "mma", "e2 parts", "epilogue"]
EVENTS = [["in_empty ok", "", "", ""], ["cv1 start", "cv1 issued", "cv2 start", "cv2 issued"], ["staging free", "converted", "", ""],
          ["e1: t1_full ok", "e1 done", "e2: t2_full ok", "e2 store issued"]]


def main():
    c = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    hw = int(sys.argv[2]) if len(sys.argv) > 2 else 320
    lib = L.require_device()
    dev = torch.device("cuda:0")
    st = L.current_stream()
    n = 64
    x = torch.randn(n, hw, hw, c, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    w1 = (torch.randn(c // 2, 3, 3, c, device=dev) / (9 * c) ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(c, 3, 3, c // 2, device=dev) / (4.5 * c) ** 0.5).to(torch.bfloat16)
    b1, b2 = torch.randn(c // 2, device=dev), torch.randn(c, device=dev)
    packed = torch.zeros((lib.dbl_bottleneck_tc_packed_bytes(c) + 15) // 16 * 4, dtype=torch.int32, device=dev)
    assert lib.dbl_bottleneck_tc_pack(w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), c, packed.data_ptr(), st) == 0
    xv, yv = L.view(x), L.view(y)
    nbytes = lib.dbl_dsconv_tc_debug_trace(None)
    buf = torch.zeros(nbytes // 8, dtype=torch.int64, device=dev)
    args = (C.byref(xv), packed.data_ptr(), 1, C.byref(yv), st)
    for _ in range(3):
        assert lib.dbl_bottleneck3x3_tc(*args) == 0, lib.dbl_last_error().decode()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.dbl_bottleneck3x3_tc(*args)
    e1.record()
    torch.cuda.synchronize()
    print(f"C={c} @{hw}^2 x{n}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    lib.dbl_dsconv_tc_debug_trace(buf.data_ptr())
    assert lib.dbl_bottleneck3x3_tc(*args) == 0
    torch.cuda.synchronize()
    lib.dbl_dsconv_tc_debug_trace(None)
    t = buf.cpu().view(4, 16, 4)
    t0 = int(t[t > 0].min())
    for it in range(8):
        print(f"tile {it}")
        for ro in range(4):
            s = "  ".join(f"{EVENTS[ro][e]}={int(t[ro, it, e]) - t0:6d}" for e in range(4) if EVENTS[ro][e] and int(t[ro, it, e]) > 0)
            print(f"   {ROLES[ro]:10s} {s}")


if __name__ == "__main__":
    main()
