"""Validator path on the B200 library: what ``models/YOLO/run.py:53`` (``model.val``) runs per batch.

Mirrors the hot part of the reference's ``DetectionValidator`` (``models/yolo/detect/val.py``):
  * ``postprocess``      - ``ops.non_max_suppression(preds, conf, iou, labels=self.lb, multi_label=True, ...)`` (val.py:92-102):
                           here the fused forward + decode + multi-label filter + a-priori labels + NMS of ``DBLBackend.detect``;
  * ``process_batch``    - ``_process_batch`` (val.py:213-227) = ``box_iou`` (utils/metrics.py:52-71) + ``match_predictions``
                           (engine/validator.py:222-260) for the WHOLE batch in one kernel (``dbl_match_predictions``), instead of a
                           per-image numpy loop after a device->host copy of every IoU matrix.
mAP bookkeeping (``ap_per_class``, plots, json) stays with the caller: it is host-side statistics over the ``tp`` matrix.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L


class DBLValidator:
    def __init__(self, model, conf: float = 0.001, iou: float = 0.7, max_det: int = 300, agnostic: bool = False):
        """model: a DBLBackend.  Defaults are the reference's validation settings (cfg/default.yaml: conf 0.001 when validating,
        iou 0.7, max_det 300)."""
        self.model, self.conf, self.iou, self.max_det, self.agnostic = model, conf, iou, max_det, agnostic
        self.iouv = torch.linspace(0.5, 0.95, 10)  # val.py:34: IoU vector for mAP@0.5:0.95
        self.niou = self.iouv.numel()

    def postprocess(self, im: torch.Tensor, labels: Optional[Sequence[torch.Tensor]] = None, labels_cap: int = 0):
        """im: BCHW batch (float 0..1 or uint8); labels: DetectionValidator.lb -- per image (n_i, 5) = (cls, x, y, w, h) in
        network-input pixels, or None (save_hybrid off, the default).  Returns (rows [B,max_det,6], counts [B])."""
        return self.model.detect(im, conf=self.conf, iou=self.iou, agnostic=self.agnostic, multi_label=True, max_det=self.max_det,
                                 labels=labels if labels is not None else None, labels_cap=labels_cap)

    def process_batch(self, rows: torch.Tensor, counts: torch.Tensor, gt_boxes: Sequence[torch.Tensor],
                      gt_cls: Sequence[torch.Tensor]) -> torch.Tensor:
        """rows/counts: NMS output in the space of gt_boxes (per image (m_i, 4) xyxy) with classes gt_cls (per image (m_i,)).
        Returns the correct-prediction matrix bool [B, max_det, niou]; image i's detections are rows [:counts[i]]."""
        return match_predictions(rows, counts, gt_boxes, gt_cls, self.iouv)


def match_predictions(rows: torch.Tensor, counts: torch.Tensor, gt_boxes: Sequence[torch.Tensor], gt_cls: Sequence[torch.Tensor],
                      iouv: torch.Tensor) -> torch.Tensor:
    """One launch of dbl_match_predictions for a batch (reference engine/validator.py:222-260 + utils/metrics.py:52-71)."""
    lib = L.require_device()
    if rows.dim() != 3 or rows.shape[2] != 6 or rows.dtype != torch.float32 or not rows.is_cuda:
        raise ValueError("rows must be a CUDA float32 [B, max_det, 6] tensor")
    B, max_det, _ = rows.shape
    if len(gt_boxes) != B or len(gt_cls) != B:
        raise ValueError(f"expected ground truth for {B} images")
    dev = rows.device
    rows = rows.contiguous()
    offs = [0]
    for g in gt_boxes:
        offs.append(offs[-1] + int(g.shape[0]))
    max_labels = max((offs[i + 1] - offs[i] for i in range(B)), default=0)
    total = offs[-1]
    boxes = torch.zeros(max(1, total), 4, dtype=torch.float32, device=dev)
    cls = torch.zeros(max(1, total), dtype=torch.float32, device=dev)
    if total:
        boxes[:total] = torch.cat([torch.as_tensor(g, dtype=torch.float32).reshape(-1, 4) for g in gt_boxes]).to(dev)
        cls[:total] = torch.cat([torch.as_tensor(c, dtype=torch.float32).reshape(-1) for c in gt_cls]).to(dev)
    offs_d = torch.tensor(offs, dtype=torch.int32, device=dev)
    thr = iouv.to(device=dev, dtype=torch.float32).contiguous()
    T = thr.numel()
    correct = torch.zeros(B, max_det, T, dtype=torch.uint8, device=dev)
    nbytes = lib.dbl_match_workspace_bytes(B, max(1, max_labels), T)
    ws = torch.empty(nbytes // 4 + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.dbl_match_predictions(rows.data_ptr(), counts.to(torch.int32).contiguous().data_ptr(), B, max_det, boxes.data_ptr(),
                                          cls.data_ptr(), offs_d.data_ptr(), max_labels, thr.data_ptr(), T, correct.data_ptr(),
                                          ws.data_ptr(), nbytes, L.current_stream()), "dbl_match_predictions")
    return correct.bool()
