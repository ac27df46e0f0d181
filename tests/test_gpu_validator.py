"""Validator path and predict tail on the GPU (-m gpu).

* a-priori labels of non_max_suppression (utils/ops.py:255-261): bit-exact against the reference's golden rows and the oracle,
  through the drop-in function and through the fused engine path (dbl_head_decode -> dbl_nms_append_labels -> dbl_nms);
* scale_boxes + clip_boxes inside the NMS epilogue (dbl_nms_scaled): bit-exact against the oracle's fp32 restatement, which is
  pinned to the reference's ops.scale_boxes byte for byte (tests/test_oracle_golden.py);
* box_iou + match_predictions (dbl_match_predictions): equal to the reference's golden `correct` matrices and the oracle.
"""
import os

import numpy as np
import pytest
import torch

import letterbox_oracle as LB
import match_oracle as M
import nms_oracle as N
from yolo_dbl_b200 import DBLBackend, DBLValidator, YOLO, match_predictions, synth_images
from yolo_dbl_b200.backend import non_max_suppression

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def _split(flat, counts):
    out, o = [], 0
    for n in counts:
        out.append(flat[o:o + n])
        o += n
    return out


@pytest.mark.parametrize("name", ["single", "multi"])
def test_nms_labels_branch_bit_exact_vs_reference_golden(name):
    g = _load(f"nms_labels_{name}.npz")
    labels = [torch.from_numpy(l) for l in _split(g["labels"], g["label_counts"].tolist())]
    out = non_max_suppression(torch.from_numpy(g["pred"]).cuda(), 0.4, 0.6, labels=labels, multi_label=bool(g["multi_label"]), max_det=300)
    assert [o.shape[0] for o in out] == g["counts"].tolist()
    assert torch.cat(out).cpu().numpy().tobytes() == g["rows"].tobytes()
    # with a class filter the label rows are filtered like any other row (ops.py:278-279)
    want, _ = N.non_max_suppression(g["pred"], 0.4, 0.6, labels=[l.numpy() for l in labels], multi_label=bool(g["multi_label"]), classes=[1, 4])
    out = non_max_suppression(torch.from_numpy(g["pred"]).cuda(), 0.4, 0.6, labels=labels, multi_label=bool(g["multi_label"]), classes=[1, 4])
    for o, w in zip(out, want):
        assert o.cpu().numpy().tobytes() == w.tobytes()


def test_validator_postprocess_with_labels_through_the_fused_engine():
    """DetectionValidator.postprocess (val.py:92-102: multi_label=True, labels=self.lb) on the production engine"""
    model = DBLBackend.synthetic("n", 3, seed=0, device="cuda:0")
    x = synth_images(3, 128, 160, 41)
    y, _ = model(x.cuda())
    g = torch.Generator().manual_seed(3)
    labels = []
    for n in (5, 0, 70):
        labels.append(torch.cat([torch.randint(0, 3, (n, 1), generator=g).float(), torch.rand(n, 2, generator=g) * 120 + 10,
                                 torch.rand(n, 2, generator=g) * 40 + 4], 1))
    val = DBLValidator(model, conf=0.001, iou=0.7)
    for lab in (labels, None):
        rows, counts = val.postprocess(x.cuda(), labels=lab)
        torch.cuda.synchronize()
        want, _ = N.non_max_suppression(y.cpu().numpy(), 0.001, 0.7, multi_label=True, max_det=300,
                                        labels=[l.numpy() for l in lab] if lab else None)
        assert counts.tolist() == [len(w) for w in want]
        for i, w in enumerate(want):
            assert rows[i, : int(counts[i])].cpu().numpy().tobytes() == w.tobytes(), i
    # a second batch of labels through the SAME captured graph (the label table is data, not graph structure)
    labels2 = [labels[2][:9], labels[0], labels[1]]
    rows, counts = val.postprocess(x.cuda(), labels=labels2, labels_cap=128)
    rows, counts = val.postprocess(x.cuda(), labels=labels, labels_cap=128)
    torch.cuda.synchronize()
    want, _ = N.non_max_suppression(y.cpu().numpy(), 0.001, 0.7, multi_label=True, max_det=300, labels=[l.numpy() for l in labels])
    assert counts.tolist() == [len(w) for w in want]


@pytest.mark.parametrize("name", ["typical", "crowded", "empty", "one"])
def test_match_predictions_equals_reference_golden(name):
    g = _load(f"match_{name}.npz")
    det, gt, gt_cls = g["det"], g["gt"], g["gt_cls"]
    # a batch of three images: the golden case, the same detections against no labels, and an image with no detections
    max_det = 300
    rows = torch.zeros(3, max_det, 6)
    rows[0, : len(det)] = torch.from_numpy(det)
    rows[1, : len(det)] = torch.from_numpy(det)
    counts = torch.tensor([len(det), len(det), 0], dtype=torch.int32)
    correct = match_predictions(rows.cuda(), counts.cuda(), [torch.from_numpy(gt), torch.zeros(0, 4), torch.from_numpy(gt)],
                                [torch.from_numpy(gt_cls), torch.zeros(0), torch.from_numpy(gt_cls)], torch.from_numpy(g["iouv"]))
    correct = correct.cpu().numpy()
    assert correct.shape == (3, max_det, 10)
    np.testing.assert_array_equal(correct[0, : len(det)], g["correct"])
    np.testing.assert_array_equal(correct[0, : len(det)], M.process_batch(det, gt, gt_cls, g["iouv"]))
    assert not correct[0, len(det):].any() and not correct[1].any() and not correct[2].any()


def test_match_predictions_random_batches_vs_oracle():
    rng = np.random.default_rng(9)
    iouv = torch.linspace(0.5, 0.95, 10)
    B, max_det = 16, 300
    rows = np.zeros((B, max_det, 6), np.float32)
    counts, gts, gcs = [], [], []
    for b in range(B):
        nd, nl = int(rng.integers(0, max_det + 1)), int(rng.integers(0, 90))
        c, s = rng.uniform(40, 600, (nl, 2)), rng.uniform(16, 120, (nl, 2))
        gt = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
        gc = rng.integers(0, 3, nl).astype(np.float32)
        for d in range(nd):
            if nl and rng.uniform() < 0.8:
                l = rng.integers(0, nl)
                rows[b, d, :4] = gt[l] + rng.normal(0, 5, 4)
                rows[b, d, 5] = gc[l] if rng.uniform() < 0.85 else rng.integers(0, 3)
            else:
                rows[b, d, :4] = np.concatenate([rng.uniform(0, 300, 2), rng.uniform(300, 640, 2)])
                rows[b, d, 5] = rng.integers(0, 3)
        rows[b, :nd, 4] = np.sort(rng.uniform(0, 1, nd))[::-1]
        counts.append(nd), gts.append(gt), gcs.append(gc)
    got = match_predictions(torch.from_numpy(rows).cuda(), torch.tensor(counts, dtype=torch.int32).cuda(), [torch.from_numpy(g) for g in gts],
                            [torch.from_numpy(c) for c in gcs], iouv).cpu().numpy()
    for b in range(B):
        want = M.process_batch(rows[b, : counts[b]], gts[b], gcs[b], iouv.numpy())
        np.testing.assert_array_equal(got[b, : counts[b]], want, err_msg=f"image {b}")
        assert not got[b, counts[b]:].any()


def _synth_bgr(h0, w0, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h0 // 8 + 2, w0 // 8 + 2, 3)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8, 1), np.float32))[:h0, :w0]
    return np.clip(img + rng.normal(0, 12, img.shape), 0, 255).astype(np.uint8)


def test_scale_boxes_in_the_nms_epilogue_bit_exact_and_results_speed():
    """YOLO.predict on image sources: boxes leave the NMS kernel in original-image coordinates (scale_boxes + clip_boxes fused,
    reference models/yolo/detect/predict.py:38); same-shape batches use the minimum-rectangle letterbox like the reference."""
    yolo = YOLO("yolov13n_DBL.yaml", nc=3)
    ragged = [_synth_bgr(h, w, 7 + i) for i, (h, w) in enumerate([(240, 320), (500, 375), (128, 96), (97, 301)])]
    same = [_synth_bgr(150, 260, 30 + i) for i in range(3)]
    for imgs, shape in ((ragged, (128, 160)), (same, (128, 128)), ([ragged[1]], (160, 160))):
        res = yolo.predict(imgs, conf=0.05, imgsz=shape)
        x = LB.preprocess(imgs, shape)  # oracle letterbox (auto rule of BasePredictor.pre_transform), pinned to the reference's cv2
        got_x = yolo.model.preprocess(imgs, shape).cpu().numpy()
        np.testing.assert_array_equal(got_x, x)
        if imgs is same:
            assert x.shape[2:] == (96, 128)  # minimum rectangle, not the 128x128 square
        rows, counts = yolo.model.detect(torch.from_numpy(x).pin_memory(), conf=0.05, iou=0.7)
        rows, counts = rows.cpu().numpy(), counts.tolist()
        assert len(res) == len(imgs) and sum(counts) > 0
        for i, (r, im) in enumerate(zip(res, imgs)):
            assert r.orig_shape == im.shape[:2] and len(r) == counts[i]
            want = rows[i, : counts[i]].copy()
            want[:, :4] = LB.scale_boxes(x.shape[2:], want[:, :4], im.shape[:2])
            assert r.boxes.data.cpu().numpy().tobytes() == want.tobytes(), i
            assert set(r.speed) == {"preprocess", "inference", "postprocess"} and all(v > 0 for v in r.speed.values())
    t = yolo.predict(synth_images(2, 64, 96, 1).cuda(), conf=0.05)
    assert set(t[0].speed) == {"preprocess", "inference", "postprocess"} and t[0].speed["inference"] > 0
