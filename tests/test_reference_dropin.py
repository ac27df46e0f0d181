"""Drop-in seams B1/B2 exercised with the REAL reference (CPU, authoring container only; skipped where /root/reference is absent).

A ``DBLBackend`` is handed to the reference's own ``AutoBackend(nn_module)`` (``nn/autobackend.py:145-155``) and
``BasePredictor.setup_model`` / ``DetectionPredictor.__call__`` (``engine/predictor.py:306-321, 219-304``).  There is no GPU here, so
the engine behind the backend is replaced by a stub that answers with the CPU oracle; everything ELSE is the product's host code
and the reference's unmodified predictor: attribute surface, ``to / fuse / float / eval / warmup``, the forward signature and
return structure, ``non_max_suppression`` on the returned tuple, ``Results`` construction.
"""
import os
import sys

import numpy as np
import pytest
import torch

import dbl_oracle as O
import ref_import
from yolo_dbl_b200 import param_shapes, resolve, synth_images, synth_state_dict
from yolo_dbl_b200.backend import DBLBackend

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="needs the reference checkout (authoring container)")


class _OracleEngine:
    """stands in for DBLEngine on a box without a GPU: same forward() contract"""

    def __init__(self, sd, scale):
        self.sd, self.scale, self.calls = sd, scale, 0

    def forward(self, x):
        self.calls += 1
        y, feats = O.forward(self.sd, x.float(), self.scale)
        return y, feats


class _CpuBackend(DBLBackend):
    def _prep(self, x):  # the product's _prep insists on an sm_100 device
        assert x.dim() == 4 and x.shape[1] == 3
        return x

    def engine(self, b, h, w, replica=0):
        if not hasattr(self, "_stub"):
            self._stub = _OracleEngine(self._sd, self.spec.scale)
        return self._stub


@pytest.fixture(scope="module")
def ref():
    ref_import.import_reference()
    import ultralytics  # noqa: F401

    return sys.modules["ultralytics"]


def test_backend_passes_through_the_reference_autobackend_and_predictor(ref):
    from ultralytics.models.yolo.detect import DetectionPredictor
    from ultralytics.nn.autobackend import AutoBackend
    from ultralytics.nn.tasks import DetectionModel, yaml_model_load

    spec = resolve(None, "n", 3)
    sd = synth_state_dict(param_shapes(spec), 0)
    backend = _CpuBackend(spec, sd, device="cpu", names={0: "algal leaf spot", 1: "brown blight", 2: "grey-blight"})
    x = synth_images(2, 128, 160, seed=21)

    # --- B1: AutoBackend(nn_module) (autobackend.py:145-155) accepts the object and forwards through it
    ab = AutoBackend(weights=backend, device=torch.device("cpu"), fp16=False, fuse=True, verbose=False)
    assert ab.pt and ab.nn_module and ab.stride == 32 and ab.names == backend.names and ab.model is backend
    ab.warmup(imgsz=(1, 3, 64, 64))
    out = ab(x)
    assert isinstance(out, list) and tuple(out[0].shape) == (2, 7, 16 * 20 + 8 * 10 + 4 * 5) and len(out[1]) == 3

    # --- B2: the reference's own predictor drives the backend end to end (tensor source -> LoadTensor -> preprocess -> model -> NMS -> Results)
    pred = DetectionPredictor(overrides=dict(conf=0.25, iou=0.7, imgsz=[128, 160], device="cpu", verbose=False, save=False))
    pred.setup_model(model=backend, verbose=False)
    assert pred.model.model is backend and pred.device.type == "cpu" and pred.args.half is False
    results = pred(source=x, stream=False)
    assert len(results) == 2 and backend._stub.calls >= 2

    # --- the same weights in the reference's own DetectionModel through the same predictor class give the same detections
    cfg = yaml_model_load("yolov13n_DBL.yaml")
    cfg["scale"] = "n"
    m = DetectionModel(cfg, ch=3, nc=3, verbose=False).eval()
    m.load_state_dict(sd, strict=True)
    m.names = backend.names
    pred2 = DetectionPredictor(overrides=dict(conf=0.25, iou=0.7, imgsz=[128, 160], device="cpu", verbose=False, save=False))
    pred2.setup_model(model=m, verbose=False)
    want = pred2(source=x, stream=False)
    n_total = 0
    for r, w in zip(results, want):
        assert r.orig_shape == w.orig_shape and r.names == w.names
        a, b = r.boxes.data.numpy(), w.boxes.data.numpy()
        assert abs(len(a) - len(b)) <= 1  # oracle (explicit BN) vs fused reference: fp32 re-association may flip a box at the threshold
        k = min(len(a), len(b))
        n_total += k
        if len(a) == len(b):
            np.testing.assert_allclose(a[:, :4], b[:, :4], atol=0.05)
            np.testing.assert_allclose(a[:, 4], b[:, 4], atol=2e-3)
            np.testing.assert_array_equal(a[:, 5], b[:, 5])
        assert set(r.speed) == {"preprocess", "inference", "postprocess"}
    assert n_total > 0
