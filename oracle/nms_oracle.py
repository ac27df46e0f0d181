"""CPU oracle for the reference's ``non_max_suppression`` — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy (strict fp32, one ufunc per arithmetic step = no fused multiply-add) restatement of
``ultralytics/utils/ops.py:167-316`` (+ ``xywh2xyxy`` ops.py:416-433) around the C restatement of
``torchvision.ops.nms`` in ``nms_oracle.c``.  Unlike the reference it also returns WHICH anchors were kept, so the
B200 path's keep-indices can be compared bit-exactly.

Deliberate deviations from the reference (documented in DESIGN.md):
  * no wall-clock abort (ops.py:238, 312-314 silently drops images when the host is slow);
  * the ``n > max_nms`` pre-truncation (ops.py:285-286) uses a STABLE descending sort (torch's CPU argsort is
    stable in practice although not requested).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "liboracle_nms.so")
        src = os.path.join(_HERE, "nms_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _LIB = ctypes.CDLL(so)
        _LIB.dbl_oracle_nms.restype = ctypes.c_int64
        _LIB.dbl_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p]
    return _LIB


def stable_desc_order(scores: np.ndarray) -> np.ndarray:
    return np.argsort(-scores.astype(np.float32), kind="stable").astype(np.int64)


def nms(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """torchvision.ops.nms(boxes[n,4] xyxy, scores[n], thr) -> kept indices (int64) in descending-score order."""
    boxes = np.ascontiguousarray(boxes, np.float32)
    n = boxes.shape[0]
    order = np.ascontiguousarray(stable_desc_order(scores))
    keep = np.empty(max(n, 1), np.int64)
    k = _lib().dbl_oracle_nms(boxes.ctypes.data, order.ctypes.data, n, ctypes.c_float(thr), keep.ctypes.data)
    return keep[:k].copy()


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """ops.py:416-433 — half = wh/2; xy-half, xy+half (each one fp32 rounding)."""
    x = x.astype(np.float32)
    half = x[..., 2:4] / np.float32(2)
    return np.concatenate([x[..., 0:2] - half, x[..., 0:2] + half], -1).astype(np.float32)


def non_max_suppression(pred: np.ndarray, conf_thres=0.25, iou_thres=0.45, classes: Optional[Sequence[int]] = None,
                        agnostic=False, multi_label=False, max_det=300, max_nms=30000, max_wh=7680, labels=None
                        ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """pred: (B, 4+nc, A) fp32 (xywh px, scores).  Returns (rows per image [n_i,6], kept anchor index per image).
    labels: per image (n_i, 5) = (cls, x, y, w, h): the validator's a-priori labels (ops.py:255-261); a kept label row has
    index A + k (k = its position in the image's label list)."""
    pred = np.asarray(pred, np.float32)
    B, no, A = pred.shape
    nc = no - 4
    multi_label = bool(multi_label) and nc > 1
    conf_thres = np.float32(conf_thres)
    out_rows, out_idx = [], []
    for b in range(B):
        x = pred[b].T  # (A, 4+nc)
        cand = np.nonzero(x[:, 4:].max(1) > conf_thres)[0]  # ops.py:234 (xc) and :253
        box = xywh2xyxy(x[cand, :4])
        cls = x[cand, 4:]
        if labels is not None and len(labels) and len(labels[b]):  # ops.py:255-261: rows appended after the filtered predictions
            lb = np.asarray(labels[b], np.float32).reshape(-1, 5)
            v = np.zeros((len(lb), nc), np.float32)
            v[np.arange(len(lb)), lb[:, 0].astype(np.int64)] = 1.0
            box = np.concatenate([box, xywh2xyxy(lb[:, 1:5])], 0)
            cls = np.concatenate([cls, v], 0)
            cand = np.concatenate([cand, A + np.arange(len(lb))])
        if multi_label:  # ops.py:270-272
            i, j = np.nonzero(cls > conf_thres)
            box, conf, cid, anchor = box[i], cls[i, j], j.astype(np.int64), cand[i]
        else:  # ops.py:273-275
            cid = cls.argmax(1) if len(cand) else np.zeros(0, np.int64)
            conf = cls[np.arange(len(cand)), cid] if len(cand) else np.zeros(0, np.float32)
            m = conf > conf_thres
            box, conf, cid, anchor = box[m], conf[m], cid[m].astype(np.int64), cand[m]
        if classes is not None:  # ops.py:278-279
            m = np.isin(cid, np.asarray(classes))
            box, conf, cid, anchor = box[m], conf[m], cid[m], anchor[m]
        if box.shape[0] > max_nms:  # ops.py:285-286
            o = stable_desc_order(conf)[:max_nms]
            box, conf, cid, anchor = box[o], conf[o], cid[o], anchor[o]
        off = cid.astype(np.float32) * np.float32(0 if agnostic else max_wh)  # ops.py:289
        keep = nms((box + off[:, None]).astype(np.float32), conf, iou_thres)[:max_det]  # ops.py:295-297
        rows = np.concatenate([box[keep], conf[keep, None], cid[keep, None].astype(np.float32)], 1).astype(np.float32)
        out_rows.append(rows.reshape(-1, 6))
        out_idx.append(anchor[keep].astype(np.int64))
    return out_rows, out_idx
