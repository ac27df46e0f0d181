"""Developer aid: time line of the roles of conv3x3_tc5_kernel's CTA 0 (producer / MMA / epilogue), first tiles.

    python tools/t5_trace.py [cin] [cout] [residual 0|1] [hw]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

ROLES = ["producer", "mma", "last quad", "epilogue"]
EVENTS = [["in_empty ok", "", "", ""], ["t_empty ok", "in_full ok", "issued", ""], ["tile done", "", "", ""],
          ["t_full ok", "in registers", "store issued", ""]]


def main():
    cin = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cout = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    res = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hw = int(sys.argv[4]) if len(sys.argv) > 4 else 80
    lib = L.require_device()
    dev = torch.device("cuda:0")
    st = L.current_stream()
    n = 64
    x = torch.randn(n, hw, hw, cin, device=dev).to(torch.bfloat16)
    y = torch.empty(n, hw, hw, cout, dtype=torch.bfloat16, device=dev)
    r = torch.randn(n, hw, hw, cout, device=dev).to(torch.bfloat16)
    wt = (torch.randn(cout, 3, 3, cin, device=dev) / (9 * cin) ** 0.5).to(torch.bfloat16)
    b = torch.randn(cout, device=dev)
    xv, yv, rv = L.view(x), L.view(y), L.view(r)
    nbytes = lib.dbl_dsconv_tc_debug_trace(None)
    buf = torch.zeros(nbytes // 8, dtype=torch.int64, device=dev)
    args = (C.byref(xv), wt.data_ptr(), b.data_ptr(), C.byref(rv) if res else None, C.byref(yv), 3, 1, 1, 1, L.ACT_SILU,
            L.EPI_ADD if res else L.EPI_NONE, L.IMPL_AUTO, st)
    assert lib.dbl_conv2d_which(C.byref(xv), C.byref(yv), 3, 1, 1, 1, L.IMPL_AUTO) == 6
    for _ in range(3):
        assert lib.dbl_conv2d(*args) == 0
    torch.cuda.synchronize()
    lib.dbl_dsconv_tc_debug_trace(buf.data_ptr())
    assert lib.dbl_conv2d(*args) == 0
    torch.cuda.synchronize()
    lib.dbl_dsconv_tc_debug_trace(None)
    t = buf.cpu().view(4, 16, 4)
    t0 = int(t[t > 0].min())
    print(f"{cin}->{cout} res={res} @{hw}^2: ns since the first stamp of CTA 0")
    for it in range(14):
        print(f"tile {it}")
        for ro in range(4):
            s = "  ".join(f"{EVENTS[ro][e]}={int(t[ro, it, e]) - t0:6d}" for e in range(4) if EVENTS[ro][e] and int(t[ro, it, e]) > 0)
            print(f"   {ROLES[ro]:10s} {s}")


if __name__ == "__main__":
    main()
