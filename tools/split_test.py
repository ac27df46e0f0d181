"""Experiment: one engine for 64 images vs S engines for 64/S images replayed concurrently on S streams."""
import sys, torch
sys.path.insert(0, '/root/repo')
from yolo_dbl_b200 import DBLBackend, synth_images
from yolo_dbl_b200.engine import DBLEngine
NMS = dict(conf=0.25, iou=0.7, max_det=300)
m = DBLBackend.synthetic("n", 80, seed=0, device="cuda:0")
x = synth_images(8, 640, 640, seed=100).repeat(8, 1, 1, 1).contiguous().cuda()
def bench(S, steps=20):
    Bp = 64 // S
    engs = [DBLEngine(m.spec, m._sd, Bp, 640, 640, device="cuda:0") for _ in range(S)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    for p, e in enumerate(engs):
        e.set_input(x[p * Bp:(p + 1) * Bp])
    torch.cuda.synchronize()
    def step():
        for e, s in zip(engs, streams):
            with torch.cuda.stream(s):
                e.detect(None, **NMS)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record()
    for s in streams: s.wait_stream(main)
    for _ in range(steps):
        step()
    for s in streams: main.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"S={S}: {ms:.3f} ms/step  {64 / ms * 1e3:.0f} images/s")
for S in (1, 2, 4):
    bench(S)
