"""CPU oracle for the image-source pre/post-process around the hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of
  * ``LetterBox.__call__`` (reference ``ultralytics/data/augment.py:1558-1600``) with ``auto=False, scaleFill=False,
    scaleup=True, center=True`` — what ``BasePredictor.pre_transform`` (``engine/predictor.py:145-161``) applies when a fixed
    network input shape is used — including the arithmetic of ``cv2.resize(..., INTER_LINEAR)`` for uint8 images;
  * ``scale_boxes`` / ``clip_boxes`` (``ultralytics/utils/ops.py:92-127, 319-338``).

Third-party dependency: ``cv2.resize`` (opencv-python, ``requirements.txt`` of the reference pins ``opencv-python>=4.6``; this
container has 4.13).  Its classic 8-bit bilinear path is fixed point: 11-bit coefficients ``cvRound(f*2048)``, a horizontal
pass in int32 and the vertical pass ``(((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2``; border samples are handled differently
along x and y (see ``_coeffs``); an exact 2x shrink takes cv2's INTER_AREA fast path, which this arithmetic reproduces.  PINNED
byte for byte against the reference's own ``LetterBox`` (which calls the installed cv2) in tests/golden/letterbox_*.npz, and
against cv2.resize on 400 random size pairs when the oracle was written.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def letterbox_params(h0: int, w0: int, H: int, W: int, auto: bool = False, stride: int = 32):
    """augment.py:1564-1597 -> (new_h, new_w, top, left, r, H_out, W_out); auto = minimum rectangle (augment.py:1577-1578)"""
    r = min(H / h0, W / w0)
    new_w, new_h = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = W - new_w, H - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_h, new_w, top, left, r, new_h + top + bottom, new_w + left + right


def _coeffs(n_out: int, n_in: int, vertical: bool = False):
    """Source indices and 11-bit weights of cv2's bilinear resize along one axis.

    cv2 treats the two axes differently at the image border (imgproc/src/resize.cpp, ``resize`` coefficient set-up vs
    ``resizeGeneric_Invoker``): along x a sample left/right of the image gets weight (2048, 0) on the edge pixel; along y the
    fractional weight is KEPT and only the two row indices are clamped, so both taps read the edge row with weights that
    each truncate in the vertical pass.  The results differ by one grey level exactly on .5 ties of border rows when upscaling.
    """
    scale = 1.0 / (n_out / n_in)  # cv2: inv_scale = dsize/ssize (double), scale = 1/inv_scale
    d = np.arange(n_out, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s).astype(np.float32)
    if not vertical:
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= n_in - 1
        f[hi], s[hi] = 0, n_in - 1
    a1 = np.rint(f * np.float32(2048)).astype(np.int32)
    a0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int32)
    return np.clip(s, 0, n_in - 1), np.clip(s + 1, 0, n_in - 1), a0, a1


def resize_linear_u8(src: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """cv2.resize(src, (new_w, new_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC images"""
    h0, w0 = src.shape[:2]
    if (h0, w0) == (new_h, new_w):
        return src.copy()
    x0, x1, ax0, ax1 = _coeffs(new_w, w0)
    y0, y1, ay0, ay1 = _coeffs(new_h, h0, vertical=True)
    s = src.astype(np.int32)
    rows = s[:, x0, :] * ax0[None, :, None] + s[:, x1, :] * ax1[None, :, None]
    r0, r1 = rows[y0], rows[y1]
    out = (((ay0[:, None, None] * (r0 >> 4)) >> 16) + ((ay1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox(img: np.ndarray, shape: Tuple[int, int], pad_value: int = 114, auto: bool = False, stride: int = 32) -> np.ndarray:
    """HWC uint8 -> HWC uint8 of `shape` (augment.py:1558-1600): resize to fit, centre, pad with 114; auto: minimum rectangle"""
    new_h, new_w, top, left, _, H, W = letterbox_params(img.shape[0], img.shape[1], shape[0], shape[1], auto, stride)
    out = np.full((H, W, img.shape[2]), pad_value, np.uint8)
    out[top:top + new_h, left:left + new_w] = resize_linear_u8(img, new_w, new_h)
    return out


def preprocess(images, shape: Tuple[int, int], auto=None) -> np.ndarray:
    """BasePredictor.preprocess for a list of HWC BGR uint8 images (engine/predictor.py:116-134, 145-161) up to, not including,
    the float cast: letterbox (auto = all images share a shape, predictor.py:155-158), BGR -> RGB, BHWC -> BCHW.  uint8 [B,3,H,W]."""
    if auto is None:
        auto = len({im.shape for im in images}) == 1
    return np.ascontiguousarray(np.stack([letterbox(im, shape, auto=auto) for im in images])[..., ::-1].transpose(0, 3, 1, 2))


def scale_boxes(img1_shape, boxes: np.ndarray, img0_shape) -> np.ndarray:
    """ops.py:92-127 (ratio_pad=None, padding=True, xyxy) followed by clip_boxes (ops.py:319-338); fp32 like the reference"""
    b = np.array(boxes, np.float32, copy=True)
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    b[..., 0] -= np.float32(pad[0])
    b[..., 1] -= np.float32(pad[1])
    b[..., 2] -= np.float32(pad[0])
    b[..., 3] -= np.float32(pad[1])
    b[..., :4] /= np.float32(gain)
    b[..., 0] = np.clip(b[..., 0], 0, img0_shape[1])
    b[..., 1] = np.clip(b[..., 1], 0, img0_shape[0])
    b[..., 2] = np.clip(b[..., 2], 0, img0_shape[1])
    b[..., 3] = np.clip(b[..., 3], 0, img0_shape[0])
    return b
