"""Developer aid: time line of the roles of dsconv_tc_kernel's CTA 0 (producer / MMA / depthwise / epilogue), first tiles.

    python tools/ds_trace.py [k] [c]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

ROLES = ["producer", "mma", "depthwise", "epilogue"]
EVENTS = [["in_empty ok", "", "", ""], ["t_empty ok", "a_full ok", "", ""], ["a_empty ok", "in_full ok", "A written", ""],
          ["t_full ok", "staging free", "store issued", ""]]


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    c = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    lib = L.require_device()
    dev = torch.device("cuda:0")
    st = L.current_stream()
    n, h, w = 64, 40, 40
    x = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
    y = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=dev)
    r = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
    wd = torch.randn(k * k, c, device=dev) / k
    wpw = (torch.randn(c, c, device=dev) / c ** 0.5).to(torch.bfloat16)
    bp = torch.randn(c, device=dev)
    xv, yv, rv = L.view(x), L.view(y), L.view(r)
    nbytes = lib.dbl_dsconv_tc_debug_trace(None)
    buf = torch.zeros(nbytes // 8, dtype=torch.int64, device=dev)
    args = (C.byref(xv), wd.data_ptr(), None, 0, wpw.data_ptr(), bp.data_ptr(), C.byref(rv), C.byref(yv), k, L.EPI_ADD, st)
    for _ in range(3):
        assert lib.dbl_dsconv_tc(*args) == 0
    torch.cuda.synchronize()
    lib.dbl_dsconv_tc_debug_trace(buf.data_ptr())
    assert lib.dbl_dsconv_tc(*args) == 0
    torch.cuda.synchronize()
    lib.dbl_dsconv_tc_debug_trace(None)
    t = buf.cpu().view(4, 16, 4)
    t0 = int(t[t > 0].min())
    print(f"k={k} c={c}: ns since the first stamp of CTA 0 (tiles of CTA 0: {-(-n * 2 * 8 // 148)})")
    for it in range(8):
        print(f"tile {it}")
        for ro in range(4):
            s = "  ".join(f"{EVENTS[ro][e]}={int(t[ro, it, e]) - t0:6d}" for e in range(4) if EVENTS[ro][e] and int(t[ro, it, e]) > 0)
            print(f"   {ROLES[ro]:10s} {s}")


if __name__ == "__main__":
    main()
