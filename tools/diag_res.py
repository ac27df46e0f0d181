import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import dbl_oracle as O
from yolo_dbl_b200 import _lib as L, param_shapes, resolve, synth_images, synth_state_dict
from yolo_dbl_b200.engine import DBLEngine
H=int(sys.argv[1]) if len(sys.argv)>1 else 1280
spec = resolve(None, "n", 3); sd = synth_state_dict(param_shapes(spec), 0)
x = synth_images(1, H, H, 6)
y_ref, feats_ref, outs_ref = O.forward(sd, x, "n", collect=True)
def rel(a,b): return float((a-b).norm()/(b.norm()+1e-12))
eng = DBLEngine(spec, sd, 1, H, H, use_graph=False, keep_layers=True)
forced = eng.debug_layers_forced(x.cuda(), outs_ref)
print("forced:", " ".join(f"{i}:{rel(o,r):.3g}" for i,(o,r) in enumerate(zip(forced[:-1], outs_ref))))
outs = eng.debug_layers(x.cuda())
print("free  :", " ".join(f"{i}:{rel(o,r):.3g}" for i,(o,r) in enumerate(zip(outs, outs_ref)) if o is not None))
y,_ = eng.forward(x.cuda()); torch.cuda.synchronize()
d=(y.cpu()-y_ref).abs(); print("layerwise engine: box p99", float(d[:,:4].flatten().quantile(0.99)), "max", float(d[:,:4].max()))
eng2 = DBLEngine(spec, sd, 1, H, H, use_graph=False)
y2,_ = eng2.forward(x.cuda()); torch.cuda.synchronize()
d=(y2.cpu()-y_ref).abs(); print("fused engine: box p99", float(d[:,:4].flatten().quantile(0.99)), "max", float(d[:,:4].max()))
A=y.shape[2]; a0=(H//8)**2; a1=a0+(H//16)**2
for nm,sl in (("P3",slice(0,a0)),("P4",slice(a0,a1)),("P5",slice(a1,A))):
    dd=(y2.cpu()-y_ref)[:,:4,sl].abs(); print(nm, "p99", float(dd.flatten().quantile(0.99)), "median", float(dd.flatten().median()))
if len(sys.argv) > 2:
    engd = DBLEngine(spec, sd, 1, H, H, use_graph=False, keep_layers=True, conv_impl=L.IMPL_DIRECT)
    forced = engd.debug_layers_forced(x.cuda(), outs_ref)
    print("forced DIRECT:", " ".join(f"{i}:{rel(o,r):.3g}" for i,(o,r) in enumerate(zip(forced[:-1], outs_ref))))
    # activation magnitude per layer in the oracle
    print("ref |max| per layer:", " ".join(f"{i}:{float(r.abs().max()):.3g}" for i,r in enumerate(outs_ref)))
