import sys, torch
sys.path.insert(0, '/root/repo')
from yolo_dbl_b200 import DBLBackend, synth_images
m = DBLBackend.synthetic("n", 80, seed=0, device="cuda:0")
eng = m.engine(64, 640, 640)
x = synth_images(8, 640, 640, seed=100).repeat(8, 1, 1, 1).contiguous()
rows, counts, _ = eng.detect(x.cuda(), conf=0.25, iou=0.7, max_det=300)
torch.cuda.synchronize()
key = eng.post_key(conf=0.25, iou=0.7, max_det=300)
c = eng._post(key)["count"].cpu()
print("candidates/img:", c[:8].tolist(), "max", int(c.max()), "mean", float(c.float().mean()))
print("kept/img:", counts[:8].tolist())
