"""Developer aid: time line of the roles of stem_tc_kernel's CTA 0 (converters / MMA / E0 / EF), first tiles.

    python tools/st_trace.py [u8 0|1] [hw]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200 import _lib as L  # noqa: E402

ROLES = ["converters", "mma", "E0 (hidden)", "EF (output)"]
EVENTS = [["in_empty ok", "tile converted", "", ""], ["stage0 start", "stage0 issued", "stage1 start", "stage1 issued"], ["t0_full ok", "hidden written", "", ""],
          ["t1_full ok", "store issued", "", ""]]


def main():
    u8 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    hw = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    lib = L.require_device()
    dev = torch.device("cuda:0")
    st = L.current_stream()
    n = 64
    x = (torch.randint(0, 256, (n, 3, hw, hw), device=dev, dtype=torch.uint8) if u8 else torch.rand(n, 3, hw, hw, device=dev))
    y = torch.empty(n, hw // 2, hw // 2, 16, dtype=torch.bfloat16, device=dev)
    w0 = torch.randn(8, 3, 3, 3, device=dev) / 27 ** 0.5
    w1 = (torch.randn(16, 3, 3, 8, device=dev) / 72 ** 0.5).to(torch.bfloat16)
    b0, b1 = torch.randn(8, device=dev), torch.randn(16, device=dev)
    packed = torch.zeros((lib.dbl_stem_tc_packed_bytes() + 15) // 16 * 4, dtype=torch.int32, device=dev)
    assert lib.dbl_stem_tc_pack(w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), b1.data_ptr(), u8, packed.data_ptr(), st) == 0
    yv = L.view(y)
    nbytes = lib.dbl_dsconv_tc_debug_trace(None)
    buf = torch.zeros(nbytes // 8, dtype=torch.int64, device=dev)
    args = (x.data_ptr(), u8, n, hw, hw, packed.data_ptr(), C.byref(yv), st)
    for _ in range(3):
        assert lib.dbl_stem_pair_tc(*args) == 0, lib.dbl_last_error().decode()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.dbl_stem_pair_tc(*args)
    e1.record()
    torch.cuda.synchronize()
    print(f"u8={u8} @{hw}^2 x{n}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    lib.dbl_dsconv_tc_debug_trace(buf.data_ptr())
    assert lib.dbl_stem_pair_tc(*args) == 0
    torch.cuda.synchronize()
    lib.dbl_dsconv_tc_debug_trace(None)
    t = buf.cpu().view(4, 16, 4)
    t0 = int(t[t > 0].min())
    for it in range(8):
        print(f"tile {it}")
        for ro in range(4):
            s = "  ".join(f"{EVENTS[ro][e]}={int(t[ro, it, e]) - t0:6d}" for e in range(4) if EVENTS[ro][e] and int(t[ro, it, e]) > 0)
            print(f"   {ROLES[ro]:12s} {s}")


if __name__ == "__main__":
    main()
