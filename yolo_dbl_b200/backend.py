"""Drop-in seams of the reference, re-hosted on libdbl_sm100.

B1  ``DBLBackend``            - a ``DetectionModel``-compatible ``nn.Module`` (reference ``nn/tasks.py:313-359``):
                                ``forward(x) -> (y, feats)``, ``.stride``, ``.names``, ``.yaml``, ``.fuse()`` ...
B2  the attribute surface ``BasePredictor`` reads off its model (reference ``engine/predictor.py:306-321``).
B3  ``non_max_suppression``   - same signature and return type as ``utils/ops.py:167-182``.
    ``YOLO`` / ``Results`` / ``Boxes`` - the thin facade the GPU box needs (the reference's own facade is unavailable
    there); same names and fields as ``engine/model.py:501-560`` and ``engine/results.py:938-1059``.

Everything here is host glue: shapes, buffers, streams.  All arithmetic runs in the CUDA library; a missing library or
GPU raises ``DblError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib as L
from . import graph as G
from .engine import DBLEngine
from .synth import synth_state_dict


def letterbox_params(h0: int, w0: int, H: int, W: int, auto: bool = False, stride: int = 32):
    """LetterBox geometry (reference data/augment.py:1564-1597; scaleup, centre): (new_h, new_w, top, left, gain, H_out, W_out).
    auto=True is the minimum-rectangle mode (augment.py:1577-1578): only the padding modulo `stride` is kept, so the output is
    the smallest stride-multiple rectangle around the resized image instead of the full H x W."""
    r = min(H / h0, W / w0)
    new_w, new_h = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = W - new_w, H - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_h, new_w, top, left, r, new_h + top + bottom, new_w + left + right


def box_scale_record(img1_shape, img0_shape):
    """(pad_x, pad_y, gain, w0, h0) of ops.scale_boxes / clip_boxes (reference utils/ops.py:92-127, 319-338; ratio_pad=None) for
    one image: what the NMS epilogue (dbl_nms_scaled) applies to every kept box."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    return [float(pad_x), float(pad_y), float(gain), float(img0_shape[1]), float(img0_shape[0])]


class DBLBackend(nn.Module):
    """B200 executor behind the reference's model interface.

    Build from a graph description + a reference-compatible ``state_dict`` (``from_state_dict``), from a live reference
    ``DetectionModel`` (``from_reference``) or with synthetic weights (``synthetic``).  Engines are specialised per input
    shape and cached.
    """

    def __init__(self, spec: G.GraphSpec, state_dict: Dict[str, torch.Tensor], device="cuda", names=None, cfg=None,
                 use_graph: bool = True, conv_impl: int = L.IMPL_AUTO):
        super().__init__()
        L.load()  # fail early and loudly when the CUDA library has not been built
        self.spec = spec
        self._sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        missing = [k for k in G.param_shapes(spec) if k not in self._sd]
        if missing:
            raise KeyError(f"state_dict lacks {len(missing)} keys, e.g. {missing[:3]}")
        self._device = torch.device(device)
        self._engines: Dict[tuple, DBLEngine] = {}
        self.use_graph, self.conv_impl = use_graph, conv_impl
        # --- attribute surface read by AutoBackend / BasePredictor
        self.stride = torch.tensor([float(s) for s in spec.strides])
        self.names = names or {i: f"{i}" for i in range(spec.nc)}
        self.nc = spec.nc
        self.yaml = cfg if isinstance(cfg, dict) else dict(G.load_cfg(cfg), scale=spec.scale, nc=spec.nc)
        self.args = {}
        self.task = "detect"
        self.pt, self.triton, self.fp16, self.dynamic, self.imx = True, False, False, False, False
        self.training = False

    # ---- construction -------------------------------------------------------------------------------------------
    @classmethod
    def from_state_dict(cls, state_dict, scale="n", nc=None, cfg=None, **kw) -> "DBLBackend":
        if nc is None:
            nc = [v.shape[0] for k, v in state_dict.items() if ".cv3." in k and k.endswith(".2.weight")][0]
        return cls(G.resolve(cfg, scale, nc), state_dict, cfg=cfg, **kw)

    @classmethod
    def from_reference(cls, model, **kw) -> "DBLBackend":
        """model: the reference's DetectionModel (fused or not; BN is folded here either way)."""
        y = model.yaml
        sd = model.state_dict()
        if not any(k.endswith(".bn.weight") for k in sd):
            raise ValueError("pass the UNFUSED reference model: fuse() drops the BatchNorm buffers the packer folds itself")
        return cls(G.resolve(y, y.get("scale", "n"), y.get("nc")), sd, names=getattr(model, "names", None), cfg=y, **kw)

    @classmethod
    def from_pt(cls, path, cfg=None, scale=None, **kw) -> "DBLBackend":
        """A reference ``.pt`` checkpoint (``torch.save({"model": DetectionModel, ...})``) without the reference package: weights, graph
        (the yaml embedded in the checkpoint), class names (``nn/tasks.py:897-924`` attempt_load_one_weight)."""
        from .checkpoint import load_pt

        sd, info = load_pt(path)
        y = cfg if cfg is not None else info["yaml"]
        sc = scale or info["scale"] or (G.scale_from_name(str(y.get("yaml_file", ""))) if isinstance(y, dict) else "n")
        return cls(G.resolve(y, sc, info["nc"]), sd, names=info["names"], cfg=y if isinstance(y, dict) else None, **kw)

    @classmethod
    def synthetic(cls, scale="n", nc=80, seed=0, cfg=None, **kw) -> "DBLBackend":
        spec = G.resolve(cfg, scale, nc)
        return cls(spec, synth_state_dict(G.param_shapes(spec), seed), cfg=cfg, **kw)

    # ---- nn.Module-ish no-ops the callers expect --------------------------------------------------------------------
    def fuse(self, verbose=False):
        return self  # BatchNorm is always folded at pack time

    def eval(self):
        return self

    def half(self):
        return self  # activations are bf16 already; the post-process stays fp32

    def float(self):
        return self

    def to(self, device=None, *a, **k):
        if device is not None and torch.device(device).type == "cuda" and torch.device(device) != self._device:
            self._device = torch.device(device)
            self._engines.clear()
        return self

    @property
    def device(self):
        return self._device

    def warmup(self, imgsz=(1, 3, 640, 640)):
        self.engine(imgsz[0], imgsz[2], imgsz[3])

    def state_dict(self, *a, **k):
        return dict(self._sd)

    # ---- execution ----------------------------------------------------------------------------------------------
    def engine(self, b: int, h: int, w: int, replica: int = 0) -> DBLEngine:
        """The engine (static buffers + captured CUDA graph) for one input shape.  ``replica`` > 0 gives further independent
        engines of the same shape: whole batches in flight on different streams (detect_stream, bench.py) fill each other's
        bubbles -- kernel fill/drain, the 64-CTA per-image kernels (hypergraph, NMS), small grids -- for ~9 % more throughput."""
        key = (int(b), int(h), int(w)) if replica == 0 else (int(b), int(h), int(w), int(replica))
        if key not in self._engines:
            self._engines[key] = DBLEngine(self.spec, self._sd, *key[:3], device=self._device, use_graph=self.use_graph,
                                           conv_impl=self.conv_impl)
        return self._engines[key]

    def _prep(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected BCHW input with 3 channels, got {tuple(x.shape)}")
        L.require_device()  # no GPU / wrong GPU -> DblError, never a silent fallback
        if x.dtype != torch.uint8 and x.dtype != torch.float32:
            x = x.float()
        return x  # host or device: the engine copies it straight into its static input buffer

    def forward(self, x, augment=False, visualize=False, embed=None):
        if augment or visualize or embed is not None:
            raise NotImplementedError("augment / visualize / embed are outside the B200 hot path")
        x = self._prep(x)
        y, feats = self.engine(*x.shape[0:1], *x.shape[2:]).forward(x)
        return y.clone(), [f.clone() for f in feats]

    __call__ = forward

    def preprocess(self, images, imgsz=(640, 640), auto=None) -> torch.Tensor:
        """Image sources: a list of HWC BGR uint8 arrays (what cv2.imread returns, any sizes) -> uint8 [B,3,H,W] RGB on the device,
        letterboxed like BasePredictor.pre_transform / LetterBox (reference engine/predictor.py:145-161, data/augment.py:1558-1600).
        One pinned upload of the raw pixels + ONE kernel (dbl_letterbox_u8) for the whole batch; the float cast and /255 happen
        inside the stem kernel.

        auto=None follows the reference's rule (predictor.py:155-158): ``LetterBox(auto=same_shapes and model.pt)`` -- when every
        image has the same shape the batch is padded only to the next stride multiple (rectangular inference), otherwise to the
        full imgsz square; True / False force a mode."""
        import numpy as np

        lib = L.require_device()
        H, W = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        if H % 32 or W % 32:
            raise ValueError("imgsz must be a multiple of 32")
        arrs = [np.ascontiguousarray(im) for im in images]
        for a in arrs:
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("expected HWC uint8 images with 3 channels (BGR)")
        same_shapes = len({a.shape for a in arrs}) == 1
        if auto is None:
            auto = same_shapes
        if auto and not same_shapes:
            raise ValueError("auto (minimum-rectangle) letterboxing needs images of one shape")
        if auto:  # the rectangle every image of the batch lands in
            H, W = letterbox_params(arrs[0].shape[0], arrs[0].shape[1], H, W, True, int(max(self.spec.strides)))[5:]
        sizes = [a.size for a in arrs]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 255) // 256 * 256)
        host = torch.empty(offs[-1], dtype=torch.uint8).pin_memory()
        for a, o in zip(arrs, offs):
            host[o:o + a.size] = torch.from_numpy(a.reshape(-1))
        with torch.cuda.device(self._device):
            raw = host.to(self._device, non_blocking=True)
            tab = (L.LbImage * len(arrs))()
            for k, (a, o) in enumerate(zip(arrs, offs)):
                h0, w0 = a.shape[:2]
                new_h, new_w, top, left = letterbox_params(h0, w0, H, W)[:4]
                tab[k] = L.LbImage(raw.data_ptr() + o, h0, w0, w0 * 3, new_h, new_w, top, left, 0)
            tab_host = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).pin_memory()
            tab_dev = tab_host.to(self._device, non_blocking=True)
            out = torch.empty(len(arrs), 3, H, W, dtype=torch.uint8, device=self._device)
            L.check(lib.dbl_letterbox_u8(tab_dev.data_ptr(), len(arrs), out.data_ptr(), H, W, 114, L.current_stream()), "dbl_letterbox_u8")
            out._dbl_keep = (raw, tab_dev)  # stream-ordered: keep the sources alive until the caller drops the result
        return out

    def detect(self, x, conf=0.25, iou=0.45, agnostic=False, multi_label=False, max_det=300, classes=None,
               max_nms=30000, orig_shapes=None, labels=None, labels_cap=0):
        """Fused forward + decode + NMS.  Returns (rows [B,max_det,6], counts [B]) living in engine-owned buffers.

        orig_shapes: per image (h0, w0) -- the kept boxes leave the NMS kernel already mapped back to the original images
        (scale_boxes + clip_boxes in the epilogue, reference models/yolo/detect/predict.py:38).
        labels: per image an (n_i, 5) tensor (cls, x, y, w, h): the validator's a-priori labels (reference utils/ops.py:255-261);
        labels_cap bounds n_i (default: the next multiple of 64 above the largest n_i)."""
        x = self._prep(x)
        eng = self.engine(x.shape[0], x.shape[2], x.shape[3])
        if labels is not None and not labels_cap:
            labels_cap = max(64, -(-max((len(l) for l in labels), default=0) // 64) * 64)
        kw = dict(conf=conf, iou=iou, agnostic=agnostic, multi_label=multi_label, max_det=max_det, classes=classes, max_nms=max_nms,
                  scaled=orig_shapes is not None, labels_cap=labels_cap if labels is not None else 0)
        key = eng.post_key(**kw)
        with torch.cuda.device(self._device):
            if orig_shapes is not None:
                rec = torch.tensor([box_scale_record(tuple(x.shape[2:]), tuple(o)) for o in orig_shapes], dtype=torch.float32)
                eng.set_scale(key, rec.pin_memory())
            if labels is not None:
                eng.set_labels(key, labels)
        rows, counts, _ = eng.detect(x, **kw)
        return rows, counts

    def detect_stream(self, batches, with_event: bool = False, in_flight: Optional[int] = None, **nms):
        """Generator over an iterable of host (pinned) or device batches: yields (rows, counts) per batch -- the analogue of
        predict(..., stream=True) (reference engine/predictor.py:219-304) -- software-pipelined three ways: the host->device
        copy of batch i+1 runs on a copy stream while batch i computes; batch i+1 is already QUEUED when batch i's result is
        handed out, so the GPU never idles while the caller reads a result; and consecutive batches run on ``in_flight``
        (default 2, env DBL_INFLIGHT) engine replicas on their own streams, so two whole batches overlap on the GPU.

        The yielded tensors are one of ``in_flight + 2`` result slots: they stay valid until the generator has been advanced TWICE.
        With ``with_event=True`` a third item is yielded, the CUDA event recorded when that batch's result was complete: a
        caller that reads results on its own stream (``stream.wait_event(ev)``) overlaps its device->host copy with the
        next batch; without it the current stream is made to wait for the batch, so plain stream order is enough."""
        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        nxt = self._prep(nxt)
        P = max(1, int(in_flight if in_flight is not None else os.environ.get("DBL_INFLIGHT", "2")))
        engs = [self.engine(nxt.shape[0], nxt.shape[2], nxt.shape[3], replica=r) for r in range(P)]
        kw = {{"agnostic_nms": "agnostic"}.get(k, k): v for k, v in nms.items()}
        with torch.cuda.device(self._device):
            caller = torch.cuda.current_stream(self._device)
            copy_stream = torch.cuda.Stream(device=self._device)
            lanes = [caller] + [torch.cuda.Stream(device=self._device) for _ in range(P - 1)]
            for ln in lanes[1:]:
                ln.wait_stream(caller)  # whatever the caller queued before (weights, earlier results) is ordered first
            S, R = P + 1, P + 2  # staging slots (one being filled + P being read), result slots (valid for two advances)
            staging = [None] * S
            ready = [None] * S
            slot_done = [None] * S
            out = [None] * R

            def launch(x, i):
                s, r, lane = i % S, i % R, lanes[i % P]
                if staging[s] is None or staging[s].dtype != x.dtype or staging[s].shape != x.shape:
                    staging[s] = torch.empty(x.shape, dtype=x.dtype, device=self._device)
                with torch.cuda.stream(copy_stream):
                    if slot_done[s] is not None:
                        copy_stream.wait_event(slot_done[s])  # the compute that read this slot has finished
                    staging[s].copy_(x, non_blocking=True)
                    ready[s] = torch.cuda.Event()
                    ready[s].record(copy_stream)
                with torch.cuda.stream(lane):
                    lane.wait_event(ready[s])
                    rows, counts, _ = engs[i % P].detect(staging[s], **kw)
                    if out[r] is None:
                        out[r] = (torch.empty_like(rows), torch.empty_like(counts))
                    out[r][0].copy_(rows, non_blocking=True)  # result slot: the engine's own buffers are rewritten by a later batch
                    out[r][1].copy_(counts, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(lane)
                slot_done[s] = ev
                return out[r][0], out[r][1], ev

            from collections import deque

            queued = deque()
            i = 0
            while nxt is not None or queued:
                while nxt is not None and len(queued) <= P:  # keep P + 1 batches queued: P computing, one whose copy is in flight
                    queued.append(launch(nxt, i))
                    i += 1
                    nxt = next(it, None)
                    if nxt is not None:
                        nxt = self._prep(nxt)
                item = queued.popleft()
                if with_event:
                    yield item
                else:
                    caller.wait_event(item[2])
                    yield item[:2]
            for ln in lanes[1:]:
                caller.wait_stream(ln)


# ---------------------------------------------------------------------------------------------------------------------
# B3: non_max_suppression
# ---------------------------------------------------------------------------------------------------------------------
class _NmsWorkspace:
    """Buffers for one (batch, capacity, max_det) shape, reused across calls."""

    _cache: Dict[tuple, "_NmsWorkspace"] = {}

    def __init__(self, dev, B, cap, max_det):
        lib = L.load()
        self.box = torch.zeros(B, cap, 4, device=dev)
        self.score = torch.zeros(B, cap, device=dev)
        self.cls = torch.zeros(B, cap, dtype=torch.int32, device=dev)
        self.pos = torch.zeros(B, cap, dtype=torch.int32, device=dev)
        self.count = torch.zeros(B, dtype=torch.int32, device=dev)
        self.keys = torch.zeros(B * lib.dbl_nms_sort_capacity(cap), dtype=torch.int64, device=dev)
        self.rows = torch.zeros(B, max_det, 6, device=dev)
        self.out_pos = torch.zeros(B, max_det, dtype=torch.int32, device=dev)
        self.out_count = torch.zeros(B, dtype=torch.int32, device=dev)
        self.cand = L.Candidates(self.box.data_ptr(), self.score.data_ptr(), self.cls.data_ptr(), self.pos.data_ptr(),
                                 self.count.data_ptr(), cap, B)

    @classmethod
    def get(cls, dev, B, cap, max_det):
        key = (str(dev), B, cap, max_det)
        if key not in cls._cache:
            cls._cache[key] = cls(dev, B, cap, max_det)
        return cls._cache[key]


def nms_raw(prediction: torch.Tensor, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
            max_det=300, max_nms=30000, max_wh=7680, labels=None):
    """Device-side result of the two NMS kernels: (rows [B,max_det,6], counts [B] int32, positions [B,max_det] int32).

    ``positions`` is the index each kept box had in the reference's filtered list ordering: the anchor index (or
    anchor*nc + class with multi_label).  No host synchronisation happens here.
    """
    lib = L.require_device()
    if prediction.dim() != 3:
        raise ValueError("prediction must be (B, 4+nc, A)")
    pred = prediction if prediction.dtype == torch.float32 else prediction.float()
    B, no, A = pred.shape
    nc = no - 4
    ml = bool(multi_label) and nc > 1
    lab = None
    if labels is not None and len(labels) and any(len(l) for l in labels):  # a-priori labels (utils/ops.py:255-261)
        if len(labels) != B:
            raise ValueError(f"labels must hold one (n_i, 5) tensor per image ({B}), got {len(labels)}")
        parts = [torch.as_tensor(l, dtype=torch.float32).reshape(-1, 5).to(pred.device) for l in labels]
        offs = [0]
        for t in parts:
            offs.append(offs[-1] + t.shape[0])
        lab = (torch.cat(parts).contiguous(), torch.tensor(offs, dtype=torch.int32, device=pred.device), max(t.shape[0] for t in parts))
    lab_cap = -(-lab[2] // 64) * 64 if lab else 0
    ws = _NmsWorkspace.get(pred.device, B, (A + lab_cap) * (nc if ml else 1), int(max_det))
    cls_t = torch.tensor(list(classes), dtype=torch.int32, device=pred.device) if classes is not None else None
    with torch.cuda.device(pred.device):
        s = L.current_stream()
        sb, sc, sa = pred.stride()
        L.check(lib.dbl_nms_filter(pred.data_ptr(), B, nc, A, sb, sc, sa, C.c_float(conf_thres), int(ml), L.ptr(cls_t),
                                   0 if cls_t is None else cls_t.numel(), C.byref(ws.cand), s), "dbl_nms_filter")
        if lab:
            L.check(lib.dbl_nms_append_labels(lab[0].data_ptr(), lab[1].data_ptr(), lab[2], A, nc, int(ml), C.c_float(conf_thres), L.ptr(cls_t),
                                              0 if cls_t is None else cls_t.numel(), C.byref(ws.cand), s), "dbl_nms_append_labels")
        L.check(lib.dbl_nms(C.byref(ws.cand), C.c_float(iou_thres), int(bool(agnostic)), C.c_float(max_wh), int(max_nms), int(max_det),
                            ws.keys.data_ptr(), ws.rows.data_ptr(), ws.out_pos.data_ptr(), ws.out_count.data_ptr(), s), "dbl_nms")
    return ws.rows, ws.out_count, ws.out_pos


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, in_place=True,
                        rotated=False) -> List[torch.Tensor]:
    """Drop-in for ``ultralytics.utils.ops.non_max_suppression`` (reference ``utils/ops.py:167-316``).

    Same arguments, same return type: a list of ``(n_i, 6)`` tensors ``(x1, y1, x2, y2, conf, cls)`` in descending score
    order.  ``labels`` (the validator's a-priori labels, ops.py:255-261) is a list of ``(n_i, 5)`` tensors
    ``(cls, x, y, w, h)``.  Differences, all deliberate: the input is never modified (``in_place`` is accepted and ignored),
    there is no wall-clock abort (``max_time_img`` ignored; reference ops.py:312-314 silently drops images), and the paths the
    YOLO-DBL head never takes (``rotated``, mask channels ``nc < C-4``) raise NotImplementedError.
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if rotated or (nc and nc != prediction.shape[1] - 4):
        raise NotImplementedError("rotated boxes and mask channels are outside the YOLO-DBL hot path")
    if prediction.shape[-1] == 6:  # end2end models (not YOLO-DBL): pass through like the reference
        return [p[p[:, 4] > conf_thres][:max_det] for p in prediction]
    rows, counts, _ = nms_raw(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms, max_wh,
                              labels=labels if labels is not None and len(labels) else None)
    counts = counts.tolist()  # the one host sync: callers of this API want variable-length tensors
    return [rows[i, : counts[i]].clone() for i in range(len(counts))]


# ---------------------------------------------------------------------------------------------------------------------
# facade: YOLO(...).predict(...) -> list[Results]
# ---------------------------------------------------------------------------------------------------------------------
class Boxes:
    """Subset of ``ultralytics.engine.results.Boxes`` (reference ``engine/results.py:938-1059``)."""

    def __init__(self, data: torch.Tensor, orig_shape):
        self.data, self.orig_shape = data, orig_shape

    xyxy = property(lambda s: s.data[:, :4])
    conf = property(lambda s: s.data[:, 4])
    cls = property(lambda s: s.data[:, 5])

    @property
    def xywh(self):
        b = self.data[:, :4]
        return torch.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], 1)

    @property
    def xyxyn(self):
        h, w = self.orig_shape
        return self.data[:, :4] / torch.tensor([w, h, w, h], device=self.data.device, dtype=self.data.dtype)

    def __len__(self):
        return self.data.shape[0]

    def cpu(self):
        return Boxes(self.data.cpu(), self.orig_shape)


class Results:
    """Subset of ``ultralytics.engine.results.Results`` (reference ``engine/results.py:187-271``)."""

    def __init__(self, boxes: torch.Tensor, orig_shape, names, path="", speed=None):
        self.boxes = Boxes(boxes, orig_shape)
        self.orig_shape, self.names, self.path = orig_shape, names, path
        self.speed = speed or {}

    def __len__(self):
        return len(self.boxes)


class YOLO:
    """``YOLO('yolov13n_DBL.yaml').predict(tensor)`` on the B200 path (reference ``models/yolo/model.py:11`` and
    ``engine/model.py:501-560``).  Only the tensor source of ``LoadTensor`` (``data/loaders.py:516-584``) is supported:
    BCHW float in 0..1 with H, W divisible by 32, so boxes need no rescaling (``scale_boxes`` is the identity)."""

    def __init__(self, model="yolov13n_DBL.yaml", task="detect", nc: Optional[int] = None, seed: int = 0, device="cuda",
                 state_dict=None, verbose=False):
        if task != "detect":
            raise NotImplementedError("only the detect task is on the YOLO-DBL hot path")
        import os

        if isinstance(model, str) and model.endswith(".pt"):  # trained / saved weights (engine/model.py:287-311 _load)
            self.model = DBLBackend.from_pt(model, device=device)
            self.names = self.model.names
            self.overrides = {"conf": 0.25, "iou": 0.7, "max_det": 300, "agnostic_nms": False, "classes": None}
            return
        scale = G.scale_from_name(os.path.basename(str(model)))
        cfg = model if (isinstance(model, dict) or (isinstance(model, str) and (os.path.exists(model) or G.builtin_cfg_name(model)))) else None
        spec = G.resolve(cfg, scale, nc)
        sd = state_dict if state_dict is not None else synth_state_dict(G.param_shapes(spec), seed)
        self.model = DBLBackend(spec, sd, device=device, cfg=cfg)
        self.names = self.model.names
        self.overrides = {"conf": 0.25, "iou": 0.7, "max_det": 300, "agnostic_nms": False, "classes": None}  # cfg/default.yaml:51-65

    def _nms_args(self, conf, iou, max_det, agnostic_nms, classes):
        o = self.overrides
        return dict(conf=o["conf"] if conf is None else conf, iou=o["iou"] if iou is None else iou,
                    agnostic=o["agnostic_nms"] if agnostic_nms is None else agnostic_nms,
                    max_det=o["max_det"] if max_det is None else max_det, classes=classes if classes is not None else o["classes"])

    def _tick(self):
        """ops.Profile (reference utils/ops.py:24-69): device-synchronised wall clock"""
        import time

        torch.cuda.synchronize(self.model.device)
        return time.perf_counter()

    def predict(self, source, conf=None, iou=None, max_det=None, agnostic_nms=None, classes=None, **kw) -> List[Results]:
        if not isinstance(source, torch.Tensor):
            return self._predict_images(source, conf, iou, max_det, agnostic_nms, classes, kw.get("imgsz", 640))
        t0 = self._tick()
        x = self.model._prep(source)
        t1 = self._tick()
        rows, counts = self.model.detect(x, **self._nms_args(conf, iou, max_det, agnostic_nms, classes))
        t2 = self._tick()
        counts = counts.tolist()
        shape = tuple(source.shape[2:])
        res = [Results(rows[i, : counts[i]].clone(), shape, self.names, path=f"image{i}.jpg") for i in range(len(counts))]
        return self._stamp(res, t0, t1, t2)

    def _stamp(self, res, t0, t1, t2):
        """Results.speed: per-image milliseconds of the three phases (reference engine/predictor.py:266-272).  Decode + NMS run
        inside the same captured graph as the forward pass, so they are counted under "inference"; "postprocess" is what is left
        of DetectionPredictor.postprocess on the host: reading the counts and wrapping the rows."""
        t3 = self._tick()
        n = max(1, len(res))
        speed = {"preprocess": (t1 - t0) * 1e3 / n, "inference": (t2 - t1) * 1e3 / n, "postprocess": (t3 - t2) * 1e3 / n}
        for r in res:
            r.speed = dict(speed)
        return res

    def _predict_images(self, source, conf, iou, max_det, agnostic_nms, classes, imgsz) -> List[Results]:
        """Image sources (one HWC BGR uint8 array or a list of them): GPU letterbox (minimum rectangle when all images share a
        shape, like the reference) -> detect, with the boxes mapped back to each original image inside the NMS epilogue
        (DetectionPredictor.postprocess, reference models/yolo/detect/predict.py:23-41)."""
        import numpy as np

        images = [source] if isinstance(source, np.ndarray) else list(source)
        if not images or not all(isinstance(im, np.ndarray) for im in images):
            raise NotImplementedError("pass a BCHW tensor, one HWC uint8 image or a list of HWC uint8 images")
        H, W = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        t0 = self._tick()
        x = self.model.preprocess(images, (H, W))
        t1 = self._tick()
        rows, counts = self.model.detect(x, orig_shapes=[im.shape[:2] for im in images], **self._nms_args(conf, iou, max_det, agnostic_nms, classes))
        t2 = self._tick()
        counts = counts.tolist()
        res = [Results(rows[i, : counts[i]].clone(), tuple(im.shape[:2]), self.names, path=f"image{i}.jpg") for i, im in enumerate(images)]
        return self._stamp(res, t0, t1, t2)

    __call__ = predict
