"""Deterministic synthetic weights ("parity weight recipe").

There is no network for checkpoints, and a freshly initialised YOLO-DBL is degenerate for
parity work (FullPAD gates are 0, DySample offsets ~1e-3, class bias ~-11 so no score ever
passes a threshold, BatchNorm statistics trivial).  This recipe fills a reference-compatible
``state_dict`` (same keys/shapes as ``ultralytics.nn.tasks.DetectionModel``, reference
``nn/tasks.py:313-359``) from ``(seed, key)`` alone, so the reference, the CPU oracle and
the B200 path can all be given bit-identical weights without shipping a checkpoint.
"""
from __future__ import annotations

import math
import os
import zlib
from typing import Dict, Mapping, Sequence

import torch

CALIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _init_pos(groups: int = 4, scale: int = 2) -> torch.Tensor:
    # value of DySample.init_pos (reference nn/modules_upsample/DySample.py:45-47): (±0.25) per sub-pixel,
    # channel = xy*groups*scale^2 + g*scale^2 + i*scale + j with x varying along j and y along i.
    v = [(-scale + 1) / 2 / scale + t / scale for t in range(scale)]
    out = torch.empty(2 * groups * scale * scale)
    for g in range(groups):
        for i in range(scale):
            for j in range(scale):
                out[g * scale * scale + i * scale + j] = v[j]
                out[groups * scale * scale + g * scale * scale + i * scale + j] = v[i]
    return out.view(1, -1, 1, 1)


def synth_tensor(key: str, shape: Sequence[int], seed: int = 0, siblings=(), recipe: str = "calibrated") -> torch.Tensor:
    g = _gen(seed, key)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    parent = key.rsplit(".", 1)[0] if "." in key else ""

    def normal(std, mean=0.0):
        return torch.randn(shape, generator=g) * std + mean

    def uniform(lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "running_mean":
        return normal(0.1)
    if leaf == "running_var":
        return uniform(0.5, 1.5)
    if leaf == "gate":
        return torch.full(shape, 0.5)
    if leaf == "gamma":  # A2C2f residual scale at the l / x scales (reference block.py:1386: 0.01 * ones; a visible value here)
        return torch.full(shape, 0.25)
    if leaf == "init_pos":
        return _init_pos().reshape(shape)
    if leaf == "prototype_base":
        return normal(0.5)
    if parent.endswith(".dfl.conv"):
        return torch.arange(shape[1], dtype=torch.float32).view(shape)
    if parent.endswith(".bn"):
        # small gammas keep SiLU near its linear range: a random deep stack with gamma ~ 1 is chaotic (it amplifies
        # bf16 rounding noise ~1.1x per layer, 50x end to end), which no trained detector is.
        if recipe == "survey":  # SURVEY.md §8d as written: weight ~ U(.5, 1.5), bias ~ N(0, .1)
            return uniform(0.5, 1.5) if leaf == "weight" else normal(0.1)
        return uniform(0.2, 0.5) if leaf == "weight" else normal(0.1)
    is_cls_out = ".cv3." in key and parent.endswith(".2")  # Detect.cv3[i][-1]
    is_box_out = ".cv2." in key and parent.endswith(".2") and len(parent.split(".")) == 5  # Detect.cv2[i][-1]
    is_lsk_out = parent.endswith(".conv") and f"{parent.rsplit('.', 1)[0]}.conv_squeeze.weight" in siblings
    if leaf == "bias":
        if is_lsk_out:  # LSKblock returns x * attn: keep attn ~ 1 + noise so stacked blocks stay O(1)
            return normal(0.1, 1.0)
        if is_cls_out:  # ~1-3 % of anchors end up above conf=0.25 (the "typical" regime of SURVEY.md §8d)
            return torch.full(shape, -(2.95 + 0.3 * math.log(shape[0])))
        if is_box_out:
            return torch.full(shape, 1.0)
        return normal(0.1)
    if leaf == "weight":
        if parent.endswith(".offset"):  # DySample.offset
            return normal(0.1)
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        if len(shape) == 2:  # nn.Linear
            return normal(1.0 / math.sqrt(fan_in))
        if is_cls_out or is_box_out:  # wide head logits: varied scores / peaky DFL distributions
            return normal((2.5 if is_cls_out else 1.0) / math.sqrt(fan_in))
        if is_lsk_out:
            return normal(0.06 / math.sqrt(fan_in))
        return normal(1.0 / math.sqrt(fan_in))
    raise KeyError(f"synth_tensor: no rule for {key} {shape}")


def synth_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0, calib="auto", recipe: str = "calibrated") -> Dict[str, torch.Tensor]:
    """shapes: {state_dict key: shape}.  Returns fp32 tensors (int64 for num_batches_tracked).

    ``calib``: BatchNorm running statistics measured on a seeded calibration batch (the analogue of what training
    would have left in the buffers; without it a random deep stack either dies out or blows up, differently at every
    width/depth).  "auto" loads ``data/bn_calib_c<C0>_nc<nc>_seed<seed>.npz`` when it exists (written by
    ``oracle/calibrate_bn.py``); a dict overrides keys directly; None keeps the recipe's N(0,.1)/U(.5,1.5) statistics.
    ``recipe="survey"`` is SURVEY.md §8d's parity recipe as written (gamma ~ U(.5,1.5), no calibration): a chaotic stack, used by
    the teacher-forced per-layer tests only.
    """
    if recipe not in ("calibrated", "survey"):
        raise ValueError(recipe)
    sd = {k: synth_tensor(k, s, seed, shapes, recipe) for k, s in shapes.items()}
    if recipe == "survey":  # the parity weight recipe of SURVEY.md §8d verbatim: BN gamma ~ U(.5,1.5), N(0,.1)/U(.5,1.5) statistics, cls bias -3
        for k in sd:
            if ".cv3." in k and k.endswith(".2.bias"):
                sd[k] = torch.full_like(sd[k], -3.0)
        return sd
    if calib == "auto":
        path = calib_path(shapes, seed)
        calib = dict(np_load(path)) if os.path.exists(path) else None
    if calib:
        for k, v in calib.items():
            if k in sd:
                sd[k] = torch.as_tensor(v, dtype=torch.float32).reshape(sd[k].shape).clone()
    return sd


def calib_path(shapes: Mapping[str, Sequence[int]], seed: int) -> str:
    c0 = shapes["model.0.conv.weight"][0]
    nc = [s[0] for k, s in shapes.items() if ".cv3." in k and k.endswith(".2.weight")][0]
    return os.path.join(CALIB_DIR, f"bn_calib_c{c0}_nc{nc}_seed{seed}.npz")


def np_load(path):
    import numpy as np

    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def synth_images(b: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """(b,3,h,w) fp32 in [0,1] with structure at every octave (smooth blobs + edges + pixel noise).

    Pure i.i.d. noise is a poor stand-in for photographs: after a few strided stages its features are spatially
    constant, BatchNorm variances collapse and the network degenerates.  Generated on the CPU so the values are
    identical on every machine.
    """
    g = torch.Generator(device="cpu").manual_seed(0x5EED ^ (seed * 7919))
    acc = torch.zeros(b, 3, h, w)
    for cell, amp in ((1, 0.35), (4, 0.5), (16, 0.8), (64, 1.0), (256, 1.0)):
        n = torch.randn(b, 3, max(2, -(-h // cell) + 1), max(2, -(-w // cell) + 1), generator=g)
        acc += amp * torch.nn.functional.interpolate(n, size=(h, w), mode="bilinear", align_corners=True)
    edges = (torch.nn.functional.interpolate(torch.randn(b, 3, max(2, h // 32), max(2, w // 32), generator=g),
                                             size=(h, w), mode="nearest") > 0.6).float()
    return torch.sigmoid(1.2 * acc + 1.5 * edges - 0.4)
