"""``.pt`` ingestion (reference nn/tasks.py:804-944) without the reference package."""
import os
import sys

import pytest
import torch

import ref_import
from ckpt_util import write_fake_reference_checkpoint
from yolo_dbl_b200 import graph as G
from yolo_dbl_b200.checkpoint import load_pt
from yolo_dbl_b200.synth import synth_state_dict


def test_load_pt_rebuilds_state_dict_yaml_and_names(tmp_path):
    spec = G.resolve(None, "n", 3)
    sd = synth_state_dict(G.param_shapes(spec), 0)
    yaml_dict = dict(G.YOLO_DBL_CFG, scale="n", nc=3, yaml_file="yolov13n_DBL.yaml")
    names = {0: "algal leaf spot", 1: "brown blight", 2: "grey-blight"}
    for use_ema in (True, False):
        path = os.path.join(tmp_path, f"fake_{use_ema}.pt")
        want = write_fake_reference_checkpoint(path, sd, yaml_dict, names, half=True, use_ema=use_ema)
        assert "ultralytics" not in sys.modules or ref_import.available()
        got, info = load_pt(path)
        assert list(got.keys()) == list(want.keys())  # same keys in the reference's registration order
        for k in want:
            assert got[k].dtype == (torch.float32 if want[k].is_floating_point() else want[k].dtype)
            assert torch.equal(got[k], want[k]), k
        assert info["scale"] == "n" and info["nc"] == 3 and info["names"] == names and info["yaml"]["backbone"][0][2] == "Conv"
        assert info["args"]["imgsz"] == 640
    # a bare state_dict file
    p2 = os.path.join(tmp_path, "sd.pt")
    torch.save(sd, p2)
    got, info = load_pt(p2)
    assert list(got.keys()) == list(sd.keys()) and info["yaml"] is None
    with pytest.raises(ValueError):
        p3 = os.path.join(tmp_path, "junk.pt")
        torch.save({"model": 3}, p3)
        load_pt(p3)


@pytest.mark.skipif(not ref_import.available(), reason="needs the reference checkout (authoring container)")
def test_load_pt_on_a_real_reference_checkpoint(tmp_path):
    """torch.save of the reference's own DetectionModel (what trainer.save_model writes), read back without its classes"""
    from copy import deepcopy

    ref_import.import_reference()
    from ultralytics.nn.tasks import DetectionModel, yaml_model_load

    cfg = yaml_model_load("yolov13n_DBL.yaml")
    m = DetectionModel(cfg, ch=3, nc=3, verbose=False)
    sd = synth_state_dict({k: list(v.shape) for k, v in m.state_dict().items()}, 0)
    m.load_state_dict(sd)
    m.names = {0: "a", 1: "b", 2: "c"}
    path = os.path.join(tmp_path, "real.pt")
    torch.save({"epoch": 1, "model": None, "ema": deepcopy(m).half(), "train_args": {"imgsz": 640}}, path)
    got, info = load_pt(path)
    ref = m.state_dict()
    assert list(got.keys()) == list(ref.keys())
    assert all(torch.equal(got[k], ref[k].half().float() if ref[k].is_floating_point() else ref[k]) for k in ref)
    assert info["scale"] == "n" and info["nc"] == 3 and info["names"] == m.names
    spec = G.resolve(info["yaml"], info["scale"], info["nc"])  # the embedded yaml drives the graph
    assert list(G.param_shapes(spec).keys()) == list(ref.keys())


@pytest.mark.gpu
def test_yolo_facade_runs_a_pt_checkpoint(tmp_path):
    import nms_oracle as N
    from yolo_dbl_b200 import DBLBackend, YOLO, synth_images

    spec = G.resolve(None, "n", 3)
    sd = synth_state_dict(G.param_shapes(spec), 0)
    path = os.path.join(tmp_path, "weights.pt")
    want_sd = write_fake_reference_checkpoint(path, sd, dict(G.YOLO_DBL_CFG, scale="n", nc=3), {0: "a", 1: "b", 2: "c"}, half=True)
    x = synth_images(2, 96, 128, 3)
    yolo = YOLO(path)
    assert yolo.names == {0: "a", 1: "b", 2: "c"}
    res = yolo.predict(x.cuda(), conf=0.1)
    direct = DBLBackend(spec, want_sd, device="cuda:0")  # the same fp16-rounded weights handed over directly
    rows, counts = direct.detect(x.cuda(), conf=0.1, iou=0.7)
    for i, r in enumerate(res):
        assert torch.equal(r.boxes.data, rows[i, : int(counts[i])])
    assert sum(len(r) for r in res) > 0
