"""Experiment: P engines of 64 images each replayed concurrently on P streams (whole batches in flight, not split ones)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from yolo_dbl_b200 import DBLBackend, synth_images
from yolo_dbl_b200.engine import DBLEngine
NMS = dict(conf=0.25, iou=0.7, max_det=300)
m = DBLBackend.synthetic("n", 80, seed=0, device="cuda:0")
x = synth_images(8, 640, 640, seed=100).repeat(8, 1, 1, 1).contiguous().cuda()
def bench(P, steps=24):
    engs = [DBLEngine(m.spec, m._sd, 64, 640, 640, device="cuda:0") for _ in range(P)]
    streams = [torch.cuda.Stream() for _ in range(P)]
    for e in engs:
        e.set_input(x)
    torch.cuda.synchronize()
    def step(i):
        with torch.cuda.stream(streams[i % P]):
            engs[i % P].detect(None, **NMS)
    for i in range(3 * P):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record()
    for s in streams: s.wait_stream(main)
    for i in range(steps):
        step(i)
    for s in streams: main.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"P={P}: {ms:.3f} ms/step  {64 / ms * 1e3:.0f} images/s", flush=True)
for P in (1, 2, 3):
    bench(P)
