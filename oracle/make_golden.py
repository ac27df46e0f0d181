"""Generate the golden fixtures under tests/golden/ by RUNNING THE REAL REFERENCE — TEST INFRASTRUCTURE.

Runs only in the authoring container (needs /root/reference; see oracle/ref_import.py).  The GPU
box never executes this; it only reads the committed ``tests/golden/*.npz|json``.

    python oracle/make_golden.py            # rewrites tests/golden/

What is recorded (all from the reference's own code, CPU fp32):
  * ``shapes_<scale>_nc<nc>.json`` state_dict key -> shape of ``DetectionModel('yolov13<scale>_DBL.yaml')``
  * ``fwd_*.npz``   y = DetectionModel.fuse().eval()(x)[0] for seeded inputs + synthetic weights, plus per-layer
                    statistics and strided samples of every one of the 36 layer outputs
  * ``nms_*.npz``   ops.non_max_suppression outputs for seeded synthetic predictions
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
from yolo_dbl_b200.synth import synth_images, synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def seeded_image(b, h, w, seed):
    return synth_images(b, h, w, seed)


def build_reference(scale, nc, seed):
    from ultralytics.nn.tasks import DetectionModel, yaml_model_load

    cfg = yaml_model_load("yolov13n_DBL.yaml")
    if scale == "l":  # BASELINE config 4: the yaml defines only n and s; inject l = [1.0, 1.0]
        cfg["scales"]["l"] = [1.0, 1.0, 114514]
    cfg["scale"] = scale
    m = DetectionModel(cfg, ch=3, nc=nc, verbose=False).eval()
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    sd = synth_state_dict(shapes, seed)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m, shapes, sd


def layer_summary(t, nsamp=64):
    flat = t.reshape(-1).double()
    idx = torch.linspace(0, flat.numel() - 1, nsamp).long()
    return np.array([flat.mean(), flat.std(), flat.abs().max()]), flat[idx].float().numpy(), idx.numpy()


def forward_fixture(name, scale, nc, b, h, w, wseed, xseed, y_stride=1, layers=True):
    m, shapes, sd = build_reference(scale, nc, wseed)
    with open(os.path.join(OUT, f"shapes_{scale}_nc{nc}.json"), "w") as f:
        json.dump(shapes, f)
    x = seeded_image(b, h, w, xseed)
    per_layer = {}
    hooks = []
    if layers:
        for i, mod in enumerate(m.model[:-1]):
            hooks.append(mod.register_forward_hook(lambda _m, _i, o, i=i: per_layer.__setitem__(i, o.detach().clone())))
    with torch.inference_mode():
        y_unfused = m(x)[0].clone()
        for hk in hooks:
            hk.remove()
        m.fuse(verbose=False)  # the predict path always fuses (nn/autobackend.py:145-155)
        y, feats = m(x)
    rec = {
        "meta": np.array(json.dumps(dict(scale=scale, nc=nc, b=b, h=h, w=w, wseed=wseed, xseed=xseed, y_stride=y_stride))),
        "y": y[:, :, ::y_stride].numpy(),
        "y_unfused": y_unfused[:, :, ::y_stride].numpy(),
    }
    for li, f in enumerate(feats):
        st, sm, ix = layer_summary(f)
        rec[f"feat{li}_stats"], rec[f"feat{li}_samples"], rec[f"feat{li}_index"] = st, sm, ix
    for i, t in per_layer.items():
        st, sm, ix = layer_summary(t)
        rec[f"layer{i}_stats"], rec[f"layer{i}_samples"], rec[f"layer{i}_index"] = st, sm, ix
        rec[f"layer{i}_shape"] = np.array(t.shape)
    np.savez_compressed(os.path.join(OUT, f"fwd_{name}.npz"), **rec)
    print(f"fwd_{name}: y{tuple(y.shape)} max score {float(y[:, 4:].max()):.3f} "
          f"cand>0.25: {float((y[:, 4:].amax(1) > 0.25).float().mean()) * 100:.2f}% "
          f"|y-y_unfused|max {float((y - y_unfused).abs().max()):.2e}")
    return m, sd, x, y


def synth_prediction(b, nc, a, seed, mu, quant=0.0, img=640.0):
    """(B, 4+nc, A) decoded prediction like Detect._inference returns: xywh px + scores in (0,1)."""
    g = torch.Generator().manual_seed(seed)
    cx = torch.rand(b, 1, a, generator=g) * img
    cy = torch.rand(b, 1, a, generator=g) * img
    wh = torch.rand(b, 2, a, generator=g) * 120 + 4
    sc = torch.sigmoid(torch.randn(b, nc, a, generator=g) * 1.5 + mu)
    if quant:
        sc = torch.round(sc / quant) * quant
    return torch.cat([cx, cy, wh, sc], 1)


def nms_fixture(name, pred_args, **kw):
    from ultralytics.utils import ops

    pred = synth_prediction(**pred_args)
    out = ops.non_max_suppression(pred.clone(), max_time_img=1e9, **kw)
    rec = {"meta": np.array(json.dumps(dict(pred_args=pred_args, kw=kw)))}
    rec["counts"] = np.array([o.shape[0] for o in out])
    rec["rows"] = torch.cat(out, 0).numpy() if len(out) else np.zeros((0, 6), np.float32)
    np.savez_compressed(os.path.join(OUT, f"nms_{name}.npz"), **rec)
    print(f"nms_{name}: kept {rec['counts'].tolist()}")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_import.import_reference()
    torch.set_num_threads(os.cpu_count())
    # --- forward fixtures ---------------------------------------------------------------
    forward_fixture("n_nc3_small", "n", 3, 2, 64, 96, wseed=0, xseed=1)
    forward_fixture("n_nc80_small", "n", 80, 1, 64, 64, wseed=3, xseed=4, layers=False)
    forward_fixture("s_nc3_small", "s", 3, 1, 64, 64, wseed=5, xseed=6, layers=False)
    forward_fixture("l_nc3_small", "l", 3, 1, 64, 64, wseed=7, xseed=8, layers=False)
    m, sd, x, y = forward_fixture("n_nc3_640", "n", 3, 1, 640, 640, wseed=0, xseed=2, y_stride=1)
    # --- NMS on the reference's own decoded output of the 640 case ------------------------
    from ultralytics.utils import ops

    for conf in (0.25, 0.001):
        out = ops.non_max_suppression(y.clone(), conf, 0.7, max_det=300, max_time_img=1e9)
        np.savez_compressed(os.path.join(OUT, f"nms_model640_conf{conf}.npz"), rows=out[0].numpy(),
                            counts=np.array([out[0].shape[0]]))
        print(f"nms_model640 conf={conf}: kept {out[0].shape[0]}")
    # --- NMS fixtures on synthetic predictions --------------------------------------------
    nms_fixture("typical_nc3", dict(b=4, nc=3, a=8400, seed=11, mu=-4.0), conf_thres=0.25, iou_thres=0.7, max_det=300)
    nms_fixture("typical_nc80", dict(b=2, nc=80, a=8400, seed=12, mu=-7.0), conf_thres=0.25, iou_thres=0.7, max_det=300)
    nms_fixture("dense_nc3", dict(b=2, nc=3, a=2100, seed=13, mu=-2.0), conf_thres=0.001, iou_thres=0.7, max_det=300)
    nms_fixture("ties_nc3", dict(b=2, nc=3, a=2100, seed=14, mu=-1.0, quant=0.125), conf_thres=0.2, iou_thres=0.5,
                max_det=300)
    nms_fixture("agnostic", dict(b=2, nc=3, a=2100, seed=15, mu=-2.0), conf_thres=0.25, iou_thres=0.45, agnostic=True,
                max_det=100)
    nms_fixture("classes", dict(b=2, nc=5, a=2100, seed=16, mu=-2.0), conf_thres=0.25, iou_thres=0.6, classes=[1, 3],
                max_det=300)
    nms_fixture("maxnms", dict(b=2, nc=3, a=2100, seed=17, mu=-1.0), conf_thres=0.05, iou_thres=0.7, max_nms=500,
                max_det=300)
    nms_fixture("empty", dict(b=2, nc=3, a=336, seed=18, mu=-9.0), conf_thres=0.25, iou_thres=0.7)
    nms_fixture("multilabel", dict(b=2, nc=3, a=2100, seed=19, mu=-2.0), conf_thres=0.1, iou_thres=0.7, multi_label=True,
                max_det=300)


if __name__ == "__main__":
    main()
