"""Micro-benchmark of the Detect post-process kernels (filter + sort/NMS) on synthetic predictions, batch 64 x 8400."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_dbl_b200.backend import nms_raw  # noqa: E402


def synth(b, nc, a, seed, mu):
    g = torch.Generator().manual_seed(seed)
    cx, cy = torch.rand(b, 1, a, generator=g) * 640, torch.rand(b, 1, a, generator=g) * 640
    wh = torch.rand(b, 2, a, generator=g) * 120 + 4
    sc = torch.sigmoid(torch.randn(b, nc, a, generator=g) * 1.5 + mu)
    return torch.cat([cx, cy, wh, sc], 1)


for name, nc, mu, conf in (("typical nc3", 3, -4.0, 0.25), ("mid nc80", 80, -5.5, 0.25), ("heavy nc3", 3, -2.0, 0.25), ("dense nc3 conf.001", 3, -1.0, 0.001)):
    pred = synth(64, nc, 8400, 5, mu).cuda()
    rows, counts, _ = nms_raw(pred, conf, 0.7)
    torch.cuda.synchronize()
    ncand = (pred[:, 4:].amax(1) > conf).sum(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        nms_raw(pred, conf, 0.7)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:22s} cand/img min {int(ncand.min())} mean {float(ncand.float().mean()):.0f} max {int(ncand.max())}  kept mean {float(counts.float().mean()):.0f}  "
          f"filter+nms {e0.elapsed_time(e1) / 10:.3f} ms/batch")
