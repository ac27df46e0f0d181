"""Round-2 golden fixtures, produced by RUNNING THE REAL REFERENCE — TEST INFRASTRUCTURE (authoring container only).

    python oracle/make_golden_r2.py

  * ``fwd_n_nc80_640.npz``            the BENCHMARKED instantiation: n-scale, nc=80, bench weights (recipe seed 0 + BN calibration
                                      table), images 2 and 3 of bench.py's batch (synth_images(8, 640, 640, seed=100)):
                                      y of DetectionModel.fuse().eval() (image 2 complete = ya, image 3 box rows + best score/class = yb_*;
                                      every 4th anchor of the UNFUSED model's y for image 2, which the oracle's explicit BatchNorm reproduces) and
                                      ops.non_max_suppression(y, 0.25, 0.7) rows for both images.
  * ``nms_labels_{single,multi}.npz`` ops.non_max_suppression(..., labels=lb) (validator path, utils/ops.py:255-261)
  * ``match_{typical,crowded,empty}.npz``  utils.metrics.box_iou + DetectionValidator.match_predictions (engine/validator.py:222-260)
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
from make_golden import build_reference, synth_prediction  # noqa: E402
from yolo_dbl_b200.synth import synth_images  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def fwd_nc80():
    from ultralytics.utils import ops

    m, shapes, sd = build_reference("n", 80, 0)
    x = synth_images(8, 640, 640, seed=100)[2:4]  # images 2 and 3 of the bench batch: the well-conditioned ones (oracle/select_bench_images.py)
    with torch.inference_mode():
        y_unfused = m(x)[0].clone()
        m.fuse(verbose=False)
        y, feats = m(x)
        out = ops.non_max_suppression(y.clone(), 0.25, 0.7, max_det=300, max_time_img=1e9)
    best, cls = y[1, 4:].max(0)
    np.savez_compressed(os.path.join(OUT, "fwd_n_nc80_640.npz"), ya=y[0].numpy(), ya_unfused_s4=y_unfused[0, :, ::4].numpy(), yb_box=y[1, :4].numpy(), yb_best=best.numpy(),
                        yb_cls=cls.numpy().astype(np.int16), rows=torch.cat(out).numpy(), counts=np.array([o.shape[0] for o in out]),
                        meta=np.array(json.dumps(dict(scale="n", nc=80, wseed=0, images="synth_images(8,640,640,seed=100)[2:4]",
                                                      conf=0.25, iou=0.7))))
    print("fwd_n_nc80_640: kept", [o.shape[0] for o in out], "cand", [(y[i, 4:].amax(0) > 0.25).sum().item() for i in range(2)])


def nms_labels():
    from ultralytics.utils import ops

    g = torch.Generator().manual_seed(77)
    for name, ml in (("single", False), ("multi", True)):
        pred = synth_prediction(b=3, nc=5, a=2100, seed=31 + ml, mu=-4.2)
        labels = []
        for n in (7, 0, 19):
            cls = torch.randint(0, 5, (n, 1), generator=g).float()
            xy = torch.rand(n, 2, generator=g) * 640
            wh = torch.rand(n, 2, generator=g) * 150 + 6
            labels.append(torch.cat([cls, xy, wh], 1))
        # two labels that coincide with strong predictions, so label rows and prediction rows suppress each other
        labels[0][0, 1:] = pred[0, :4, 5]
        labels[2][3, 1:] = pred[2, :4, 17]
        out = ops.non_max_suppression(pred.clone(), 0.4, 0.6, labels=labels, multi_label=ml, max_det=300, max_time_img=1e9)
        np.savez_compressed(os.path.join(OUT, f"nms_labels_{name}.npz"), pred=pred.numpy(), rows=torch.cat(out).numpy(),
                            counts=np.array([o.shape[0] for o in out]), labels=torch.cat(labels).numpy(),
                            label_counts=np.array([len(l) for l in labels]), multi_label=np.array(ml))
        print(f"nms_labels_{name}: kept", [o.shape[0] for o in out])


def match_cases():
    from ultralytics.engine.validator import BaseValidator
    from ultralytics.utils.metrics import box_iou

    iouv = torch.linspace(0.5, 0.95, 10)
    stub = types.SimpleNamespace(iouv=iouv)
    rng = np.random.default_rng(5)
    for name, (nd, nl, jitter) in {"typical": (120, 14, 6.0), "crowded": (300, 60, 12.0), "empty": (40, 0, 0.0), "one": (1, 1, 1.0)}.items():
        gt_xy = rng.uniform(40, 600, (nl, 2))
        gt_wh = rng.uniform(20, 160, (nl, 2))
        gt = np.concatenate([gt_xy - gt_wh / 2, gt_xy + gt_wh / 2], 1).astype(np.float32)
        gt_cls = rng.integers(0, 4, nl).astype(np.float32)
        det = np.zeros((nd, 6), np.float32)
        for d in range(nd):
            if nl and rng.uniform() < 0.7:  # a detection near some label, sometimes with the wrong class
                l = rng.integers(0, nl)
                det[d, :4] = gt[l] + rng.normal(0, jitter, 4)
                det[d, 5] = gt_cls[l] if rng.uniform() < 0.8 else rng.integers(0, 4)
            else:
                c, s = rng.uniform(40, 600, 2), rng.uniform(20, 160, 2)
                det[d, :4] = np.concatenate([c - s / 2, c + s / 2])
                det[d, 5] = rng.integers(0, 4)
        det[:, 4] = np.sort(rng.uniform(0.001, 1, nd))[::-1]
        dt, gtt, gct = torch.from_numpy(det), torch.from_numpy(gt), torch.from_numpy(gt_cls)
        if nl:
            iou = box_iou(gtt, dt[:, :4])
            correct = BaseValidator.match_predictions(stub, dt[:, 5], gct, iou).numpy()
        else:
            iou, correct = torch.zeros(0, nd), np.zeros((nd, 10), bool)
        np.savez_compressed(os.path.join(OUT, f"match_{name}.npz"), det=det, gt=gt, gt_cls=gt_cls, iou=iou.numpy(), correct=correct,
                            iouv=iouv.numpy())
        print(f"match_{name}: tp@0.5 {int(correct[:, 0].sum())} tp@0.95 {int(correct[:, -1].sum())} of {nd} detections / {nl} labels")


def fwd_sibling(name, yaml_name, seed, xseed, hw):
    """a sibling graph of the repository on the synthetic recipe: the reference's fused / unfused outputs + per-layer samples"""
    from ultralytics.nn.tasks import DetectionModel, yaml_model_load
    from yolo_dbl_b200.synth import synth_state_dict

    cfg = yaml_model_load(yaml_name)
    m = DetectionModel(cfg, ch=3, nc=3, verbose=False).eval()
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, f"shapes_{name}.json"), "w") as f:
        json.dump(shapes, f)
    sd = synth_state_dict(shapes, seed)
    m.load_state_dict(sd, strict=True)
    x = synth_images(1, hw[0], hw[1], seed=xseed)
    per_layer = {}
    hooks = [mod.register_forward_hook(lambda _m, _i, o, i=i: per_layer.__setitem__(i, o.detach().clone())) for i, mod in enumerate(m.model[:-1])]
    with torch.inference_mode():
        y_unfused = m(x)[0].clone()
        for h in hooks:
            h.remove()
        m.fuse(verbose=False)
        y = m(x)[0]
    rec = dict(y=y.numpy(), y_unfused=y_unfused.numpy())
    for i, t in per_layer.items():
        flat = t.reshape(-1)
        idx = torch.linspace(0, flat.numel() - 1, 64).long()
        rec[f"layer{i}_samples"], rec[f"layer{i}_index"], rec[f"layer{i}_shape"] = flat[idx].numpy(), idx.numpy(), np.array(t.shape)
    np.savez_compressed(os.path.join(OUT, f"fwd_{name}_small.npz"), **rec)
    print(f"fwd_{name}_small: y", tuple(y.shape), "max score", float(y[:, 4:].max()), "|fused-unfused|", float((y - y_unfused).abs().max()))


def fwd_dbl2():
    """yolov13l_DBL2.yaml (C3Ghost / GhostBottleneck / DSC3k2 variants): the reference's output on the synthetic recipe (seed 9)"""
    from ultralytics.nn.tasks import DetectionModel, yaml_model_load
    from yolo_dbl_b200.synth import synth_state_dict

    cfg = yaml_model_load("yolov13l_DBL2.yaml")
    m = DetectionModel(cfg, ch=3, nc=3, verbose=False).eval()
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, "shapes_l_DBL2_nc3.json"), "w") as f:
        json.dump(shapes, f)
    sd = synth_state_dict(shapes, 9)
    m.load_state_dict(sd, strict=True)
    x = synth_images(1, 64, 96, seed=10)
    per_layer = {}
    hooks = [mod.register_forward_hook(lambda _m, _i, o, i=i: per_layer.__setitem__(i, o.detach().clone())) for i, mod in enumerate(m.model[:-1])]
    with torch.inference_mode():
        y_unfused = m(x)[0].clone()
        for h in hooks:
            h.remove()
        m.fuse(verbose=False)
        y = m(x)[0]
    rec = dict(y=y.numpy(), y_unfused=y_unfused.numpy())
    for i, t in per_layer.items():
        flat = t.reshape(-1)
        idx = torch.linspace(0, flat.numel() - 1, 64).long()
        rec[f"layer{i}_samples"], rec[f"layer{i}_index"], rec[f"layer{i}_shape"] = flat[idx].numpy(), idx.numpy(), np.array(t.shape)
    np.savez_compressed(os.path.join(OUT, "fwd_l_DBL2_nc3_small.npz"), **rec)
    print("fwd_l_DBL2_nc3_small: y", tuple(y.shape), "max score", float(y[:, 4:].max()), "|fused-unfused|", float((y - y_unfused).abs().max()))


if __name__ == "__main__":
    ref_import.import_reference()
    torch.set_num_threads(os.cpu_count())
    match_cases()
    nms_labels()
    fwd_nc80()
    fwd_dbl2()
    fwd_sibling("n_V13_nc3", "yolov13n.yaml", 13, 14, (128, 160))  # stock yolov13.yaml: A2C2f area attention (area 4 and 1), grouped convs
