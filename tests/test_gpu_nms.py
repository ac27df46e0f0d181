"""Detect post-process parity (-m gpu).

* decode: dbl_detect_decode vs the oracle's Detect._inference restatement on the same head logits, 1e-3 relative.
* NMS: fed the REFERENCE's decoded boxes (golden y) and seeded synthetic predictions, the kept rows must equal the
  reference's / the oracle's bit for bit (keep-indices included), incl. ties, class filter, agnostic, max_nms
  truncation, multi_label, empty images, and a full-size B=64 x 8400 batch.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import dbl_oracle as O
import nms_oracle as N
from yolo_dbl_b200 import _lib as L
from yolo_dbl_b200.backend import nms_raw, non_max_suppression

from conftest import GOLDEN
from gpu_util import dev, stream

pytestmark = pytest.mark.gpu


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def _synth_prediction(b, nc, a, seed, mu, quant=0.0, img=640.0):
    g = torch.Generator().manual_seed(seed)
    cx = torch.rand(b, 1, a, generator=g) * img
    cy = torch.rand(b, 1, a, generator=g) * img
    wh = torch.rand(b, 2, a, generator=g) * 120 + 4
    sc = torch.sigmoid(torch.randn(b, nc, a, generator=g) * 1.5 + mu)
    if quant:
        sc = torch.round(sc / quant) * quant
    return torch.cat([cx, cy, wh, sc], 1)


NMS_CASES = ["typical_nc3", "typical_nc80", "dense_nc3", "ties_nc3", "agnostic", "classes", "maxnms", "empty", "multilabel"]


@pytest.mark.parametrize("name", NMS_CASES)
def test_nms_bit_exact_vs_reference_golden(name):
    g = _load(f"nms_{name}.npz")
    meta = json.loads(str(g["meta"]))
    pred = _synth_prediction(**meta["pred_args"])
    out = non_max_suppression(pred.to(dev()), **meta["kw"])
    assert [o.shape[0] for o in out] == g["counts"].tolist()
    got = torch.cat(out, 0).cpu().numpy() if out else np.zeros((0, 6), np.float32)
    assert got.astype(np.float32).tobytes() == g["rows"].astype(np.float32).tobytes()
    # keep indices vs the oracle (the reference does not expose them)
    _, idx = N.non_max_suppression(pred.numpy(), **meta["kw"])
    rows, counts, pos = nms_raw(pred.to(dev()), meta["kw"].get("conf_thres", 0.25), meta["kw"].get("iou_thres", 0.45),
                                meta["kw"].get("classes"), meta["kw"].get("agnostic", False), meta["kw"].get("multi_label", False),
                                meta["kw"].get("max_det", 300), meta["kw"].get("max_nms", 30000))
    nc = meta["pred_args"]["nc"]
    for b, want in enumerate(idx):
        got_pos = pos[b, : int(counts[b])].cpu().numpy().astype(np.int64)
        if meta["kw"].get("multi_label"):
            got_pos = got_pos // nc
        np.testing.assert_array_equal(got_pos, want)


def test_nms_on_reference_decoded_boxes():
    y = torch.from_numpy(_load("fwd_n_nc3_640.npz")["y"])
    for conf in (0.25, 0.001):
        ref = _load(f"nms_model640_conf{conf}.npz")["rows"].astype(np.float32)
        out = non_max_suppression(y.to(dev()), conf, 0.7, max_det=300)
        assert out[0].cpu().numpy().tobytes() == ref.tobytes()


def test_nms_input_not_modified_and_tuple_accepted():
    pred = _synth_prediction(2, 3, 500, 1, -2.0).to(dev())
    keep = pred.clone()
    a = non_max_suppression((pred, None), 0.25, 0.5)
    assert torch.equal(pred, keep)
    b = non_max_suppression(pred, 0.25, 0.5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_nms_full_size_batch64_matches_oracle():
    # BASELINE config 2: 64 x (4+nc) x 8400, typical and dense regimes
    for nc, mu, conf in ((3, -4.0, 0.25), (3, -1.0, 0.001)):
        pred = _synth_prediction(64, nc, 8400, 100 + nc, mu)
        rows, counts, pos = nms_raw(pred.to(dev()), conf, 0.7)
        torch.cuda.synchronize()
        sel = [0, 17, 63]
        want_rows, want_idx = N.non_max_suppression(pred[sel].numpy(), conf, 0.7)
        for j, b in enumerate(sel):
            n = int(counts[b])
            assert n == len(want_idx[j])
            assert rows[b, :n].cpu().numpy().tobytes() == want_rows[j].tobytes()
            np.testing.assert_array_equal(pos[b, :n].cpu().numpy().astype(np.int64), want_idx[j])
        # size-independent properties on the whole batch: scores sorted, counts bounded, every kept row is a candidate
        c = counts.cpu()
        assert int(c.max()) <= 300 and int(c.min()) >= 0
        r = rows.cpu()
        for b in range(64):
            s = r[b, : int(c[b]), 4]
            assert bool((s[:-1] >= s[1:]).all()) and bool((s > conf).all())


@pytest.mark.parametrize("nc,hw", [(5, [(16, 24), (8, 12), (4, 6)]), (80, [(40, 40), (20, 20), (10, 10)]), (3, [(12, 20), (6, 10), (3, 5)])])
def test_decode_vs_oracle_and_fused_filter_consistency(nc, hw):
    lib = L.require_device()
    g = torch.Generator().manual_seed(7)
    B = 3
    strides = [8.0, 16.0, 32.0]
    pitch = (64 + nc + 3) // 4 * 4
    feats = [torch.cat([torch.randn(B, 64, h, w, generator=g) * 2, torch.randn(B, nc, h, w, generator=g) * 1.5 - 1.0], 1) for h, w in hw]
    A = sum(h * w for h, w in hw)
    ref = O.detect_decode(torch.cat([f.reshape(B, 64 + nc, -1) for f in feats], 2), hw, strides)
    bufs = []
    lv = L.HeadLevels()
    for i, f in enumerate(feats):
        buf = torch.zeros(B, hw[i][0], hw[i][1], pitch, device=dev())
        buf[..., : 64 + nc] = f.permute(0, 2, 3, 1).to(dev())
        bufs.append(buf)
        lv.data[i], lv.h[i], lv.w[i], lv.stride[i] = buf.data_ptr(), hw[i][0], hw[i][1], strides[i]
    lv.num_levels, lv.batch, lv.nc, lv.pitch = 3, B, nc, pitch
    y = torch.zeros(B, 4 + nc, A, device=dev())
    assert lib.dbl_detect_decode(C.byref(lv), y.data_ptr(), stream()) == 0, lib.dbl_last_error().decode()
    torch.cuda.synchronize()
    torch.testing.assert_close(y.cpu(), ref, rtol=1e-3, atol=1e-3)  # north-star tolerance for the fp32 post-process
    # fused decode+filter -> nms must equal decode -> filter -> nms exactly (same arithmetic, no y round trip)
    for ml in (False, True):
        rows_a, cnt_a, pos_a = [t.clone() for t in nms_raw(y, 0.25, 0.6, multi_label=ml)]
        cap = A * (nc if ml else 1)
        ws = dict(box=torch.zeros(B, cap, 4, device=dev()), score=torch.zeros(B, cap, device=dev()),
                  cls=torch.zeros(B, cap, dtype=torch.int32, device=dev()), pos=torch.zeros(B, cap, dtype=torch.int32, device=dev()),
                  count=torch.zeros(B, dtype=torch.int32, device=dev()))
        cand = L.Candidates(ws["box"].data_ptr(), ws["score"].data_ptr(), ws["cls"].data_ptr(), ws["pos"].data_ptr(),
                            ws["count"].data_ptr(), cap, B)
        keys = torch.zeros(B * lib.dbl_nms_sort_capacity(cap), dtype=torch.int64, device=dev())
        rows = torch.zeros(B, 300, 6, device=dev())
        opos = torch.zeros(B, 300, dtype=torch.int32, device=dev())
        ocnt = torch.zeros(B, dtype=torch.int32, device=dev())
        assert lib.dbl_decode_filter(C.byref(lv), C.c_float(0.25), int(ml), None, 0, C.byref(cand), stream()) == 0
        assert lib.dbl_nms(C.byref(cand), C.c_float(0.6), 0, C.c_float(7680.0), 30000, 300, keys.data_ptr(), rows.data_ptr(),
                           opos.data_ptr(), ocnt.data_ptr(), stream()) == 0
        torch.cuda.synchronize()
        assert torch.equal(ocnt, cnt_a) and torch.equal(rows, rows_a) and torch.equal(opos, pos_a)
        # and the whole chain equals the oracle's NMS of the CUDA-decoded y
        want, _ = N.non_max_suppression(y.cpu().numpy(), 0.25, 0.6, multi_label=ml)
        for b in range(B):
            assert rows[b, : int(ocnt[b])].cpu().numpy().tobytes() == want[b].tobytes()
