"""Parity on the configuration bench.py measures (-m gpu): YOLO-DBL-n, nc=80, 640x640, B=64, the PRODUCTION engine -- fused stem
(dbl_stem_pair), fused Detect tail (dbl_head_decode<80,80>), CUDA graph with 8 lanes, two engine replicas through detect_stream --
with bench.py's own weights (DBLBackend.synthetic("n", 80, seed=0)) and images (synth_images(8, 640, 640, seed=100) tiled to 64).

  * forward(): y against the fp32 oracle on the 8 distinct images and against the REAL reference's output (golden
    fwd_n_nc80_640.npz, images 2 and 3 of the batch).  The synthetic network is ill-conditioned on some of the bench images (the
    oracle's own forward with bf16 activation storage moves boxes by up to 55 px p99 there), so the absolute
    bounds are asserted on the well-conditioned images and every image is held to a multiple of the oracle's own bf16 drift.
  * detect(): rows and counts bit-exact against the oracle's non_max_suppression of the CUDA y -- images 0 (3 365 candidates,
    the slow NMS path), 31, 63 and the rest; float and uint8 input; multi_label; class filter.
Bounds marked MEASURED are <= 2x the value printed by this test on B200 (round 2).
"""
import os

import numpy as np
import pytest
import torch

import dbl_oracle as O
import nms_oracle as N
from yolo_dbl_b200 import DBLBackend, synth_images

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

B, HW = 64, 640
BOX_MED, BOX_P99, SCORE_P99 = 0.032, 0.8, 7.5e-3  # absolute bounds on the well-conditioned images = 2 x MEASURED (0.016 px, 0.386 px, 0.0037)
WELL = (2, 3, 6)  # images of the 8 whose fp32 reference is well-conditioned (oracle bf16 drift: box p99 < 0.3 px, score p99 < 2e-3)


@pytest.fixture(scope="module")
def bench():
    model = DBLBackend.synthetic("n", 80, seed=0, device="cuda:0")
    small = synth_images(8, HW, HW, seed=100)
    x = small.repeat(B // 8, 1, 1, 1).contiguous()  # exactly bench.py's batch
    eng = model.engine(B, HW, HW)
    y, feats = eng.forward(x.pin_memory())
    y = y.clone()
    torch.cuda.synchronize()
    y_ref, _ = O.forward(model.state_dict(), small, "n")
    y_b16, _ = O.forward(model.state_dict(), small, "n", round_bf16=True)
    return dict(model=model, eng=eng, small=small, x=x, y=y, y_ref=y_ref, y_b16=y_b16)


def _stats(d):
    box, sc = d[:4].flatten(), d[4:].flatten()
    return float(box.median()), float(box.quantile(0.99)), float(box.max()), float(sc[::5].quantile(0.99)), float(sc.max())


def test_production_engine_is_the_benchmarked_instantiation(bench):
    eng = bench["eng"]
    assert eng.fuse_stem and eng.fuse_head and eng.use_graph and eng.n_lanes == 8
    assert eng.head_fused is not None and (eng.head_fused.c2, eng.head_fused.c3, eng.head_fused.nc) == (64, 80, 80)
    kerns = {m["kern"] for m in eng.op_meta}
    assert {"stem_pair", "bottleneck_tc", "conv_s2_tc", "conv_tc5", "conv_tc4", "conv_tc_1x1", "dsconv_tc"} <= kerns, kerns


def test_forward_vs_oracle_and_reference(bench):
    y, y_ref, y_b16 = bench["y"].cpu(), bench["y_ref"], bench["y_b16"]
    assert bool(torch.isfinite(y).all())
    # the 8 distinct images repeat 8 times in the batch: every copy must give the same answer (per-image arithmetic does not
    # depend on the position in the batch)
    for rep in range(1, B // 8):
        assert torch.equal(y[rep * 8:(rep + 1) * 8], y[:8]), f"batch positions {rep * 8}.. differ from 0..7"
    stats = [(_stats((y[i] - y_ref[i]).abs()), _stats((y_b16[i] - y_ref[i]).abs())) for i in range(8)]
    for i, (ours, drift) in enumerate(stats):
        print(f"image {i}: CUDA vs fp32 oracle  box med {ours[0]:.3f} p99 {ours[1]:.3f} max {ours[2]:.2f} px | score p99 {ours[3]:.4f} max {ours[4]:.4f}"
              f"   || oracle-bf16 vs fp32: box med {drift[0]:.3f} p99 {drift[1]:.3f} max {drift[2]:.2f} | score p99 {drift[3]:.4f} max {drift[4]:.4f}")
    for i, (ours, drift) in enumerate(stats):
        # every image, chaotic or not: the bulk of the boxes (median) is as accurate as the reference's own bf16 run allows
        assert ours[0] <= 3.0 * drift[0] + 0.02, f"image {i}: median box error {ours[0]:.3f} vs bf16 drift {drift[0]:.3f}"
        if drift[1] < 1.0:  # the fp32 reference is not chaotic on this image (its own bf16 run keeps p99 < 1 px): tails are comparable too
            assert ours[1] <= 3.0 * drift[1] + 0.1, f"image {i}: p99 box error {ours[1]:.3f} vs bf16 drift {drift[1]:.3f}"
            assert ours[3] <= 3.0 * drift[3] + 1e-3, f"image {i}: p99 score error {ours[3]:.4f} vs bf16 drift {drift[3]:.4f}"
        if i in WELL:  # absolute bounds where the reference is well-conditioned
            assert ours[0] <= BOX_MED and ours[1] <= BOX_P99 and ours[3] <= SCORE_P99, (i, ours)
    assert sum(1 for _, d in stats if d[1] < 1.0) >= 5  # the tail comparison above really ran on most of the batch
    # the REAL reference (fused DetectionModel) on images 2 and 3 of the batch
    g = np.load(os.path.join(GOLDEN, "fwd_n_nc80_640.npz"))
    s = _stats((y[2] - torch.from_numpy(g["ya"])).abs())
    print(f"image 2 vs REAL reference: box med {s[0]:.3f} p99 {s[1]:.3f} max {s[2]:.2f} px | score p99 {s[3]:.4f} max {s[4]:.4f}")
    assert s[0] <= BOX_MED and s[1] <= BOX_P99 and s[3] <= SCORE_P99, s
    d3 = (y[3, :4] - torch.from_numpy(g["yb_box"])).abs().flatten()
    best = y[3, 4:].max(0).values
    db = float((best - torch.from_numpy(g["yb_best"])).abs().quantile(0.99))
    print(f"image 3 vs REAL reference: box p99 {float(d3.quantile(0.99)):.3f} px | best-class score p99 {db:.4f}")
    assert float(d3.quantile(0.99)) <= BOX_P99 and db <= 1.5e-2  # best-class scores are the large ones: measured 0.0075, bound 2x


@pytest.mark.parametrize("dtype", ["f32", "u8"])
def test_detect_rows_bit_exact_vs_oracle_nms_of_cuda_y(bench, dtype):
    eng = bench["eng"]
    x = bench["x"] if dtype == "f32" else (bench["x"] * 255.0).round().to(torch.uint8)
    y, _ = eng.forward(x.pin_memory())  # fused head with logits + y outputs (same kernel instantiation as detect())
    y = y.cpu().numpy()
    for kw in (dict(conf=0.25, iou=0.7), dict(conf=0.25, iou=0.7, multi_label=True), dict(conf=0.3, iou=0.5, classes=[0, 17, 79]),
               dict(conf=0.05, iou=0.6, agnostic=True, max_det=100)):
        rows, counts, pos = eng.detect(x.pin_memory(), **kw)
        torch.cuda.synchronize()
        rows, counts = rows.cpu().numpy(), counts.tolist()
        images = (0, 1, 2, 3, 4, 5, 6, 7, 31, 63) if kw == dict(conf=0.25, iou=0.7) else (0, 31, 63)
        okw = dict(max_det=kw.get("max_det", 300), multi_label=kw.get("multi_label", False), classes=kw.get("classes"), agnostic=kw.get("agnostic", False))
        want, _ = N.non_max_suppression(y[list(images)], kw["conf"], kw["iou"], **okw)
        for i, w in zip(images, want):
            assert counts[i] == len(w), (dtype, kw, i, counts[i], len(w))
            assert rows[i, : counts[i]].tobytes() == w.tobytes(), (dtype, kw, i)
        if kw == dict(conf=0.25, iou=0.7):
            print(f"[{dtype}] detections per 64-image step: {sum(counts)} (bench.py reports this as detections_per_step)")
            assert counts[:8] * 8 == counts


def test_two_replicas_in_flight_equal_single_engine(bench):
    """bench.py's device leg and detect_stream run consecutive steps on two engine replicas: same bits as one engine"""
    model, x = bench["model"], bench["x"]
    xu = (x * 255.0).round().to(torch.uint8).pin_memory()
    rows, counts = model.detect(xu, conf=0.25, iou=0.7)
    want = (rows.clone(), counts.clone())
    n = 0
    for r, c in model.detect_stream([xu] * 4, in_flight=2, conf=0.25, iou=0.7):
        assert torch.equal(c, want[1]) and torch.equal(r, want[0])
        n += 1
    assert n == 4


def test_s_scale_nc80_production_engine_head_instantiation():
    """dbl_head_decode<128,80> (s scale: c3 = 128): fused tail vs the layer-wise engine and the oracle"""
    from yolo_dbl_b200 import param_shapes, resolve, synth_state_dict
    from yolo_dbl_b200.engine import DBLEngine

    spec = resolve(None, "s", 80)
    sd = synth_state_dict(param_shapes(spec), 5)
    x = synth_images(2, 160, 192, 12)
    e_prod = DBLEngine(spec, sd, 2, 160, 192)
    e_layer = DBLEngine(spec, sd, 2, 160, 192, use_graph=False, keep_layers=True)
    assert e_prod.head_fused is not None and e_prod.head_fused.c3 == 128 and e_layer.head_fused is None
    y1 = e_prod.forward(x.cuda())[0].clone()
    y2 = e_layer.forward(x.cuda())[0].clone()
    torch.cuda.synchronize()
    d = (y1 - y2).abs()
    print(f"s/80 fused vs layer-wise: box p99 {float(d[:, :4].flatten().quantile(.99)):.3f} max {float(d[:, :4].max()):.3f} score max {float(d[:, 4:].max()):.4f}")
    assert float(d[:, :4].flatten().quantile(0.99)) <= 1.0 and float(d[:, 4:].max()) <= 3e-2
    for kw in (dict(conf=0.25, iou=0.7), dict(conf=0.1, iou=0.6, multi_label=True), dict(conf=0.1, iou=0.6, classes=[3, 40])):
        rows, counts, _ = e_prod.detect(x.cuda(), **kw)
        torch.cuda.synchronize()
        want, _ = N.non_max_suppression(y1.cpu().numpy(), kw["conf"], kw["iou"], max_det=300, multi_label=kw.get("multi_label", False), classes=kw.get("classes"))
        assert counts.tolist() == [len(w) for w in want], kw
        for i, w in enumerate(want):
            assert rows[i, : int(counts[i])].cpu().numpy().tobytes() == w.tobytes(), (kw, i)
