"""CPU oracle for the validator's matching step — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy (strict fp32, one ufunc per step) restatement of
  * ``box_iou``            ``ultralytics/utils/metrics.py:52-71``
  * ``match_predictions``  ``ultralytics/engine/validator.py:222-260`` (``use_scipy=False``), called from
                           ``DetectionValidator._process_batch`` (``models/yolo/detect/val.py:213-227``).
Pinned by tests/golden/match_*.npz, produced by calling the reference's own two functions (oracle/make_golden_r2.py).
Pairs of exactly equal IoU: the reference orders them with an unstable argsort; this restatement (like the CUDA kernel) lets
the pair with the larger label index win, which is what numpy's argsort()[::-1] yields for short lists.
"""
from __future__ import annotations

import numpy as np


def box_iou(box1: np.ndarray, box2: np.ndarray, eps: float = 1e-7) -> np.ndarray:
    """(N,4) x (M,4) xyxy -> (N,M) fp32"""
    a = np.asarray(box1, np.float32)[:, None, :]
    b = np.asarray(box2, np.float32)[None, :, :]
    wh = np.clip(np.minimum(a[..., 2:], b[..., 2:]) - np.maximum(a[..., :2], b[..., :2]), 0, None).astype(np.float32)
    inter = wh[..., 0] * wh[..., 1]
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    return (inter / (((area_a + area_b) - inter) + np.float32(eps))).astype(np.float32)


def match_predictions(pred_classes: np.ndarray, true_classes: np.ndarray, iou: np.ndarray, iouv: np.ndarray) -> np.ndarray:
    """iou: (L labels, D detections).  Returns correct bool (D, len(iouv)) -- the reference's sequence of numpy steps."""
    pred_classes, true_classes = np.asarray(pred_classes, np.float32), np.asarray(true_classes, np.float32)
    correct = np.zeros((pred_classes.shape[0], len(iouv)), bool)
    correct_class = true_classes[:, None] == pred_classes[None, :]
    iou = (np.asarray(iou, np.float32) * correct_class.astype(np.float32)).astype(np.float32)
    for i, threshold in enumerate(np.asarray(iouv, np.float32)):
        matches = np.array(np.nonzero(iou >= threshold)).T
        if matches.shape[0]:
            if matches.shape[0] > 1:
                order = np.argsort(iou[matches[:, 0], matches[:, 1]], kind="stable")[::-1]  # ties: the later pair first
                matches = matches[order]
                matches = matches[np.unique(matches[:, 1], return_index=True)[1]]
                matches = matches[np.unique(matches[:, 0], return_index=True)[1]]
            correct[matches[:, 1].astype(int), i] = True
    return correct


def process_batch(rows: np.ndarray, gt_boxes: np.ndarray, gt_cls: np.ndarray, iouv: np.ndarray) -> np.ndarray:
    """DetectionValidator._process_batch: rows (D,6) detections, gt (L,4)/(L,) -> correct (D, T)"""
    if len(gt_cls) == 0 or len(rows) == 0:
        return np.zeros((len(rows), len(iouv)), bool)
    return match_predictions(rows[:, 5], gt_cls, box_iou(gt_boxes, rows[:, :4]), iouv)
