"""Whole-graph parity (-m gpu): the B200 path vs the CPU oracle on identical weights and inputs.

Stated tolerance for the bf16 backbone (north star: "bf16 backbone within a stated tolerance"):
  * teacher-forced, every top-level layer (inputs = the oracle's outputs rounded to bf16): relative L2 error of the
    layer's output <= 1.5e-2 (2 x the largest measured value, 0.007 at the DSC3k2 layer; bf16 rounding of every intermediate inside
    the block, fp32 accumulation); head logits <= 1.5e-2;
    also run with SURVEY.md §8d's weight recipe exactly as written (BN gamma ~ U(.5, 1.5), uncalibrated);
  * free-running: every layer's relative L2 error is bounded by a multiple of what bf16 ACTIVATION STORAGE costs the reference
    itself on the same weights, input and shape -- the oracle run with every block output rounded to bf16
    (O.forward(round_bf16=True)) against its fp32 run: err_cuda(layer) <= 2 * err_bf16_oracle(layer) + 3e-3 (measured ratio <= 1.65).  No constant is
    invented per scale: a chaotic synthetic stack gets a loose bound exactly as far as the reference's own bf16 run is loose;
  * decoded outputs at 640x640 (n scale): the bounds in test_full_640_* are MEASURED values x 2 (printed by the test).
The fp32 post-process on identical inputs is held to 1e-3 / bit-exactness in test_gpu_nms.py.
"""
import json
import os

import numpy as np
import pytest
import torch

import dbl_oracle as O
import nms_oracle as N
from yolo_dbl_b200 import DBLBackend, YOLO, _lib as L, param_shapes, resolve, synth_images, synth_state_dict
from yolo_dbl_b200.engine import DBLEngine

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

# decoded-output bounds of the n-scale 640x640 free-running case = 2 x the values measured on B200 (printed by the test)
BOX_MED_640, BOX_P99_640, BOX_MAX_640, SCORE_P99_640, SCORE_MAX_640 = 0.022, 0.22, 1.4, 4e-3, 4.7e-2  # measured 0.011 / 0.108 / 0.700 px, 0.0020 / 0.0232


def _rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def _check_layers(scale, nc, seed, b, h, w, conv_impl, xseed=3, forced_tol=1.5e-2, recipe="calibrated", free=True):
    spec = resolve(None, scale, nc)
    sd = synth_state_dict(param_shapes(spec), seed, recipe=recipe)
    x = synth_images(b, h, w, xseed)
    eng = DBLEngine(spec, sd, b, h, w, device="cuda:0", conv_impl=conv_impl, use_graph=False, keep_layers=True)
    y_ref, feats_ref, outs_ref = O.forward(sd, x, scale, collect=True)
    _, _, outs_b16 = O.forward(sd, x, scale, collect=True, round_bf16=True)  # the reference's own cost of bf16 activation storage
    head_ref = torch.cat([f.reshape(b, f.shape[1], -1) for f in feats_ref], 2)
    # --- teacher-forced: each layer on the oracle's inputs
    forced = eng.debug_layers_forced(x.cuda(), outs_ref)
    errs = []
    for i, (o, r) in enumerate(zip(forced, outs_ref + [head_ref])):
        assert o.shape == r.shape, f"layer {i}: shape {tuple(o.shape)} vs {tuple(r.shape)}"
        assert bool(torch.isfinite(o).all()), f"layer {i} produced non-finite values"
        errs.append((_rel_l2(o, r), i, spec.layers[i].kind))
    print(f"[{scale} {h}x{w} impl={conv_impl}] teacher-forced rel-L2:", " ".join(f"{i}:{e:.3g}" for e, i, _ in errs))
    for e, i, kind in errs:
        assert e <= forced_tol, f"teacher-forced layer {i} ({kind}) relative L2 error {e:.3g} (conv_impl={conv_impl})"
    # --- free-running: bounded by the reference's own bf16 drift at the same layer
    outs = eng.debug_layers(x.cuda())
    fr = [(_rel_l2(o, r), _rel_l2(q, r), i, spec.layers[i].kind) for i, (o, r, q) in enumerate(zip(outs, outs_ref, outs_b16)) if o is not None]
    print(f"[{scale} {h}x{w} impl={conv_impl}] free-running  rel-L2 (ours/oracle-bf16):", " ".join(f"{i}:{e:.3g}/{q:.3g}" for e, q, i, _ in fr))
    for e, q, i, kind in fr:
        if free:
            assert e <= 2.0 * q + 3e-3, f"free-running layer {i} ({kind}): error {e:.3g} vs the reference's own bf16 drift {q:.3g} (conv_impl={conv_impl})"
    y, feats = eng.forward(x.cuda())
    torch.cuda.synchronize()
    return eng, sd, x, y.cpu(), y_ref


@pytest.mark.parametrize("conv_impl", [L.IMPL_DIRECT, L.IMPL_AUTO])
def test_layers_n_scale_small(conv_impl):
    _check_layers("n", 3, 0, 2, 128, 160, conv_impl)


def test_layers_n_scale_nc80():
    _check_layers("n", 80, 3, 1, 96, 64, L.IMPL_AUTO)


def test_layers_s_scale():
    _check_layers("s", 3, 5, 1, 96, 96, L.IMPL_AUTO)


def test_layers_l_scale():
    _check_layers("l", 3, 7, 1, 64, 64, L.IMPL_AUTO)


def test_layers_survey_recipe_as_written():
    """SURVEY.md §8d's parity weight recipe verbatim (BN gamma ~ U(.5, 1.5), N(0,.1) / U(.5,1.5) statistics, no calibration table)"""
    _check_layers("n", 3, 11, 2, 128, 160, L.IMPL_AUTO, recipe="survey")
    _check_layers("s", 3, 12, 1, 96, 128, L.IMPL_AUTO, recipe="survey")


def test_full_640_vs_oracle_and_reference_golden():
    eng, sd, x, y, y_ref = _check_layers("n", 3, 0, 1, 640, 640, L.IMPL_AUTO, xseed=2)
    g = np.load(os.path.join(GOLDEN, "fwd_n_nc3_640.npz"))
    y_gold = torch.from_numpy(g["y"])  # the REAL reference's output on the same weights/input
    for name, ref in (("oracle", y_ref), ("reference", y_gold)):
        dbox = (y[:, :4] - ref[:, :4]).abs().flatten()
        dsc = (y[:, 4:] - ref[:, 4:]).abs().flatten()
        print(f"vs {name}: box median {dbox.median():.3f} p99 {dbox.quantile(0.99):.3f} max {dbox.max():.3f} px | "
              f"score p99 {dsc.quantile(0.99):.4f} max {dsc.max():.4f}")
        # MEASURED x 2 (B200, round 2)
        assert float(dbox.median()) <= BOX_MED_640 and float(dbox.quantile(0.99)) <= BOX_P99_640 and float(dbox.max()) <= BOX_MAX_640
        assert float(dsc.quantile(0.99)) <= SCORE_P99_640 and float(dsc.max()) <= SCORE_MAX_640
    # detections vs the reference's own NMS output: nearly every reference box has a twin (IoU > 0.9, same class)
    rows, counts, _ = eng.detect(x.cuda(), conf=0.25, iou=0.7)
    ref_rows = torch.from_numpy(np.load(os.path.join(GOLDEN, "nms_model640_conf0.25.npz"))["rows"])
    got = rows[0, : int(counts[0])].cpu()
    import torchvision

    iou = torchvision.ops.box_iou(ref_rows[:, :4], got[:, :4])
    same = ref_rows[:, 5:6] == got[:, 5].unsqueeze(0)
    matched = ((iou > 0.9) & same).any(1).float().mean()
    print(f"detections: reference {len(ref_rows)}, ours {len(got)}, matched {float(matched) * 100:.1f}%")
    assert abs(len(got) - len(ref_rows)) <= max(3, len(ref_rows) // 10) and float(matched) >= 0.9


def test_cuda_graph_replay_equals_eager():
    spec = resolve(None, "n", 3)
    sd = synth_state_dict(param_shapes(spec), 0)
    x = synth_images(2, 96, 128, 9).cuda()
    e1 = DBLEngine(spec, sd, 2, 96, 128, use_graph=False)
    e2 = DBLEngine(spec, sd, 2, 96, 128, use_graph=True)
    y1 = e1.forward(x)[0].clone()
    for _ in range(3):
        y2 = e2.forward(x)[0]
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    r1 = [t.clone() for t in e1.detect(x, conf=0.1, iou=0.7)]
    r2 = [t.clone() for t in e2.detect(x, conf=0.1, iou=0.7)]
    r2 = [t.clone() for t in e2.detect(x, conf=0.1, iou=0.7)]
    assert all(torch.equal(a, b) for a, b in zip(r1, r2))


def test_backend_api_surface_and_facade():
    m = DBLBackend.synthetic("n", 3, seed=0, device="cuda:0")
    assert m.stride.tolist() == [8.0, 16.0, 32.0] and m.pt and not m.triton and m.task == "detect"
    assert m.fuse() is m and m.eval() is m and m.half() is m
    x = synth_images(2, 64, 96, 1)
    y, feats = m(x.cuda())
    assert tuple(y.shape) == (2, 7, 8 * 12 + 4 * 6 + 2 * 3) and [tuple(f.shape) for f in feats] == [(2, 67, 8, 12), (2, 67, 4, 6), (2, 67, 2, 3)]
    # y and feats are consistent with each other through the oracle's decode
    hw = [f.shape[2:] for f in feats]
    y2 = O.detect_decode(torch.cat([f.cpu().reshape(2, 67, -1) for f in feats], 2), hw, [8.0, 16.0, 32.0])
    torch.testing.assert_close(y.cpu(), y2, rtol=1e-3, atol=1e-3)
    res = YOLO("yolov13n_DBL.yaml", nc=3).predict(x.cuda(), conf=0.05)
    assert len(res) == 2 and res[0].boxes.data.shape[1] == 6 and res[0].orig_shape == (64, 96)
    assert bool((res[0].boxes.conf[:-1] >= res[0].boxes.conf[1:]).all()) if len(res[0]) > 1 else True


def test_uint8_input_and_pipelined_stream_match_float_path():
    m = DBLBackend.synthetic("n", 3, seed=0, device="cuda:0")
    xs = [(synth_images(2, 96, 128, 20 + i) * 255).round().to(torch.uint8) for i in range(3)]
    # uint8 (0..255 exact in bf16, the /255 folded into layer 0's weights) vs float(x)/255 (pixels rounded to bf16 by the
    # stem): the two differ by one bf16 rounding of layer 0's operands -> boxes within 0.25 px, scores within 5e-3
    y_f, _ = m(xs[0].float() / 255.0)
    y_u, _ = m(xs[0].pin_memory())
    torch.cuda.synchronize()
    d = (y_u - y_f).abs()
    assert float(d[:, :4].max()) <= 0.25 and float(d[:, 4:].max()) <= 5e-3, (float(d[:, :4].max()), float(d[:, 4:].max()))
    # one call per batch from pinned host memory ...
    want = []
    for x in xs:
        rows, counts = m.detect(x.pin_memory(), conf=0.1, iou=0.7)
        want.append((rows.clone(), counts.clone()))
    # ... equals the pipelined generator (copy of batch i+1 overlaps compute of batch i) bit for bit
    got = [(r.clone(), c.clone()) for r, c in m.detect_stream([x.pin_memory() for x in xs], conf=0.1, iou=0.7)]
    assert len(got) == 3
    for (r, c), (wr, wc) in zip(got, want):
        assert torch.equal(c, wc) and torch.equal(r, wr)
    assert int(sum(int(c.sum()) for _, c in got)) > 0
    # any number of engine replicas in flight (own buffers, own graph, own stream) gives the same detections, with and without
    # the completion event; five batches so that staging / result slots and replicas are all reused
    xs5 = xs + [(synth_images(2, 96, 128, 30 + i) * 255).round().to(torch.uint8) for i in range(2)]
    want5 = []
    for x in xs5:
        rows, counts = m.detect(x.pin_memory(), conf=0.1, iou=0.7)
        want5.append((rows.clone(), counts.clone()))
    for p in (1, 2, 3):
        got = [(r.clone(), c.clone()) for r, c in m.detect_stream([x.pin_memory() for x in xs5], in_flight=p, conf=0.1, iou=0.7)]
        assert len(got) == 5
        for (r, c), (wr, wc) in zip(got, want5):
            assert torch.equal(c, wc) and torch.equal(r, wr), f"in_flight={p}"
    side = torch.cuda.Stream()
    n = 0
    for (r, c, ev), (wr, wc) in zip(m.detect_stream([x.pin_memory() for x in xs5], with_event=True, in_flight=2, conf=0.1, iou=0.7), want5):
        side.wait_event(ev)
        with torch.cuda.stream(side):
            rc, cc = r.cpu(), c.cpu()  # synchronous copies on the side stream: the caller has read the slot before advancing
        assert torch.equal(cc, wc.cpu()) and torch.equal(rc, wr.cpu())
        n += 1
    assert n == 5


def test_fused_stem_engine_matches_layerwise_engine():
    """The production engine fuses layers 0+1 (dbl_stem_pair); the layer-wise engine used by the parity tests does not."""
    spec = resolve(None, "n", 3)
    sd = synth_state_dict(param_shapes(spec), 0)
    x = synth_images(2, 128, 160, 4)
    e_layer = DBLEngine(spec, sd, 2, 128, 160, use_graph=False, keep_layers=True)
    e_fused = DBLEngine(spec, sd, 2, 128, 160, use_graph=False)
    assert not e_layer.fuse_stem and e_fused.fuse_stem
    assert any(m["kern"] == "stem_pair" for m in e_fused.op_meta) and any(m["kern"] == "bottleneck_tc" for m in e_fused.op_meta)
    y_ref, _ = O.forward(sd, x, "n")
    for name, inp in (("f32", x.cuda()), ("u8", (x * 255).round().to(torch.uint8).cuda())):
        y1 = e_layer.forward(inp)[0].clone()
        y2 = e_fused.forward(inp)[0].clone()
        torch.cuda.synchronize()
        d12 = (y1 - y2).abs()
        dref = (y2.cpu() - y_ref).abs()
        print(f"[{name}] fused vs layer-wise: box max {float(d12[:, :4].max()):.3f} px, score max {float(d12[:, 4:].max()):.4f}; "
              f"fused vs oracle: box p99 {float(dref[:, :4].flatten().quantile(0.99)):.3f} px")
        assert float(d12[:, :4].flatten().quantile(0.99)) <= 1.5 and float(d12[:, 4:].max()) <= 5e-2
        if name == "f32":
            assert float(dref[:, :4].flatten().quantile(0.99)) <= 1.5 and float(dref[:, 4:].flatten().quantile(0.99)) <= 1e-2
    # fused Detect tail (dbl_head_decode): feats, y and the candidate lists come from the same kernel and must agree
    assert e_fused.head_fused is not None and e_layer.head_fused is None
    y2, feats = e_fused.forward(x.cuda())
    torch.cuda.synchronize()
    hw = [tuple(f.shape[2:]) for f in feats]
    y_from_feats = O.detect_decode(torch.cat([f.cpu().reshape(2, f.shape[1], -1) for f in feats], 2), hw, [8.0, 16.0, 32.0])
    torch.testing.assert_close(y2.cpu(), y_from_feats, rtol=1e-3, atol=1e-3)
    for kw in (dict(conf=0.25, iou=0.7), dict(conf=0.05, iou=0.5, multi_label=True), dict(conf=0.1, iou=0.6, classes=[0, 2])):
        rows, counts, _ = e_fused.detect(x.cuda(), **kw)
        torch.cuda.synchronize()
        want, _ = N.non_max_suppression(y2.cpu().numpy(), kw["conf"], kw["iou"], max_det=300, multi_label=kw.get("multi_label", False),
                                        classes=kw.get("classes"))
        assert counts.tolist() == [len(w) for w in want], (kw, counts.tolist(), [len(w) for w in want])
        for i, w in enumerate(want):
            assert rows[i, : int(counts[i])].cpu().numpy().tobytes() == w.tobytes(), (kw, i)


def test_config5_1280_dense_nms_through_the_production_engine():
    """BASELINE config 5: n-scale at 1280^2 (33 600 anchors > max_nms = 30 000), validator threshold conf = 0.001."""
    spec = resolve(None, "n", 3)
    sd = synth_state_dict(param_shapes(spec), 0)
    x = synth_images(1, 1280, 1280, 6)
    eng = DBLEngine(spec, sd, 1, 1280, 1280)
    y, _ = eng.forward(x.cuda())
    y = y.clone()
    torch.cuda.synchronize()
    assert tuple(y.shape) == (1, 7, 33600) and bool(torch.isfinite(y).all())
    # free-running bound: a multiple of the reference's own bf16 drift on this input (at 1280^2 the synthetic network is in its
    # high-gain regime, |x| ~ 60: the oracle's bf16 run itself moves); what this config is about -- 33 600 anchors, max_nms
    # truncation, dense NMS -- is checked bit-exactly below.
    y_ref, _ = O.forward(sd, x, "n")
    y_b16, _ = O.forward(sd, x, "n", round_bf16=True)
    rel, rel_b16 = _rel_l2(y.cpu(), y_ref), _rel_l2(y_b16, y_ref)
    print(f"1280^2: rel-L2 of y {rel:.4f} (oracle bf16 drift {rel_b16:.4f})")
    assert rel <= 2.0 * rel_b16 + 3e-3
    for kw in (dict(conf=0.001, iou=0.7), dict(conf=0.001, iou=0.7, multi_label=True)):
        rows, counts, _ = eng.detect(x.cuda(), **kw)
        torch.cuda.synchronize()
        want, _ = N.non_max_suppression(y.cpu().numpy(), kw["conf"], kw["iou"], max_det=300, multi_label=kw.get("multi_label", False))
        assert counts.tolist() == [len(w) for w in want]
        assert rows[0, : int(counts[0])].cpu().numpy().tobytes() == want[0].tobytes(), kw


def test_config4_l_scale_production_engine():
    """BASELINE config 4: the l width/depth of the same yaml (44.7 M parameters), production (fused) engine vs the oracle."""
    spec = resolve(None, "l", 3)
    sd = synth_state_dict(param_shapes(spec), 7)
    x = synth_images(2, 192, 256, 8)
    eng = DBLEngine(spec, sd, 2, 192, 256)
    y, _ = eng.forward(x.cuda())
    y = y.clone()
    torch.cuda.synchronize()
    y_ref, _ = O.forward(sd, x, "l")
    assert bool(torch.isfinite(y).all())
    y_b16, _ = O.forward(sd, x, "l", round_bf16=True)
    rel, rel_b16 = _rel_l2(y.cpu(), y_ref), _rel_l2(y_b16, y_ref)
    print(f"l scale 192x256: rel-L2 of y {rel:.4f} (oracle bf16 drift {rel_b16:.4f})")
    assert rel <= 2.0 * rel_b16 + 3e-3
    rows, counts, _ = eng.detect(x.cuda(), conf=0.25, iou=0.7)
    want, _ = N.non_max_suppression(y.cpu().numpy(), 0.25, 0.7, max_det=300)
    assert counts.tolist() == [len(w) for w in want]


def _synth_bgr(h0, w0, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h0 // 8 + 2, w0 // 8 + 2, 3)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8, 1), np.float32))[:h0, :w0]
    return np.clip(img + rng.normal(0, 12, img.shape), 0, 255).astype(np.uint8)


def test_letterbox_kernel_bit_exact_vs_oracle_and_reference_golden():
    """dbl_letterbox_u8: integer work -> bit-exact against the oracle's fixed-point restatement of
    LetterBox/cv2.resize AND against the reference's own cv2 output (golden fixtures made by running the reference)."""
    import letterbox_oracle as LB

    m = DBLBackend.synthetic("n", 3, seed=0, device="cuda:0")
    # ragged batch: down/up-scaling, identity, odd sizes, a 1-pixel-wide strip, in one launch
    sizes = [(150, 260), (50, 37), (96, 128), (97, 131), (301, 299), (480, 640), (1080, 1920), (17, 400), (3, 5)]
    imgs = [_synth_bgr(h, w, 100 + i) for i, (h, w) in enumerate(sizes)]
    for shape in [(96, 128), (160, 160), (640, 640)]:
        got = m.preprocess(imgs, shape).cpu().numpy()  # ragged batch -> full-size letterbox (auto off), like the reference
        want = LB.preprocess(imgs, shape)
        assert got.shape == want.shape and got.dtype == np.uint8
        np.testing.assert_array_equal(got, want)
    for name in ["down_wide", "up_tall", "same", "odd", "square_down", "up2x", "down2x", "up_both", "thin", "auto_wide", "auto_tall"]:
        g = np.load(os.path.join(GOLDEN, f"letterbox_{name}.npz"))
        shape = tuple(int(v) for v in g["shape"])
        got = m.preprocess([g["img"]], shape, auto=name.startswith("auto_")).cpu().numpy()[0]
        ref = np.ascontiguousarray(g["out"][..., ::-1].transpose(2, 0, 1))
        np.testing.assert_array_equal(got, ref)


def test_config4_l_scale_640_batch_and_config5_1280_batch16():
    """BASELINE configs 4 and 5 at their own shapes: l width/depth at 640x640 (B = 4; the bench runs B = 32 of the same engine
    shape class) and n at 1280x1280 with B = 16, conf 0.001.  y against the oracle on the distinct images, bounded by the
    reference's own bf16 drift; detections bit-exact against the oracle's NMS of the CUDA y."""
    for scale, seed, b, hw, xseed, conf in (("l", 7, 4, 640, 8, 0.25), ("n", 0, 16, 1280, 6, 0.001)):
        spec = resolve(None, scale, 3)
        sd = synth_state_dict(param_shapes(spec), seed)
        small = synth_images(2, hw, hw, xseed)
        x = small.repeat(b // 2, 1, 1, 1).contiguous()
        eng = DBLEngine(spec, sd, b, hw, hw)
        y = eng.forward(x.pin_memory())[0].clone()
        torch.cuda.synchronize()
        assert bool(torch.isfinite(y).all()) and torch.equal(y[:2], y[b - 2:])
        y_ref, _ = O.forward(sd, small, scale)
        y_b16, _ = O.forward(sd, small, scale, round_bf16=True)
        rel, rel_b16 = _rel_l2(y[:2].cpu(), y_ref), _rel_l2(y_b16, y_ref)
        d = (y[:2].cpu() - y_ref).abs()
        print(f"{scale} {hw}^2 B={b}: rel-L2 of y {rel:.4f} (oracle bf16 drift {rel_b16:.4f}); box p99 {float(d[:, :4].flatten().quantile(.99)):.3f} px")
        assert rel <= 2.0 * rel_b16 + 3e-3
        rows, counts, _ = eng.detect(x.pin_memory(), conf=conf, iou=0.7)
        torch.cuda.synchronize()
        want, _ = N.non_max_suppression(y[:2].cpu().numpy(), conf, 0.7, max_det=300)
        for i, w in enumerate(want):
            for j in (i, b - 2 + i):
                assert int(counts[j]) == len(w) and rows[j, : len(w)].cpu().numpy().tobytes() == w.tobytes(), (scale, j)
        del eng
        torch.cuda.empty_cache()


def test_sibling_graph_dbl2_c3ghost():
    """yolov13l_DBL2.yaml: C3Ghost / GhostBottleneck / GhostConv and the DSC3k2 variants on the B200 path, teacher-forced
    per layer and free-running against the oracle (which is pinned to the reference's own DBL2 output, tests/test_oracle_golden.py)."""
    spec = resolve("yolov13l_DBL2.yaml", "l", 3)
    sd = synth_state_dict(param_shapes(spec), 9)
    x = synth_images(1, 64, 96, 10)
    eng = DBLEngine(spec, sd, 1, 64, 96, use_graph=False, keep_layers=True)
    y_ref, feats_ref, outs_ref = O.forward(sd, x, "l", collect=True, cfg="DBL2")
    _, _, outs_b16 = O.forward(sd, x, "l", collect=True, round_bf16=True, cfg="DBL2")
    head_ref = torch.cat([f.reshape(1, f.shape[1], -1) for f in feats_ref], 2)
    forced = eng.debug_layers_forced(x.cuda(), outs_ref)
    errs = [(_rel_l2(o, r), i, spec.layers[i].kind) for i, (o, r) in enumerate(zip(forced, outs_ref + [head_ref]))]
    print("[DBL2 l 64x96] teacher-forced rel-L2:", " ".join(f"{i}:{e:.3g}" for e, i, _ in errs))
    for e, i, kind in errs:
        assert e <= 1.5e-2, f"teacher-forced layer {i} ({kind}) relative L2 error {e:.3g}"
    outs = eng.debug_layers(x.cuda())
    for i, (o, r, q) in enumerate(zip(outs, outs_ref, outs_b16)):
        if o is not None:
            assert _rel_l2(o, r) <= 2.0 * _rel_l2(q, r) + 3e-3, f"free-running layer {i} ({spec.layers[i].kind})"
    # production engine (graph, fused head) through the facade, detections bit-exact vs the oracle NMS of its own y
    m = DBLBackend.synthetic("l", 3, seed=9, cfg="yolov13l_DBL2.yaml", device="cuda:0")
    y, _ = m(x.cuda())
    rows, counts = m.detect(x.cuda(), conf=0.01, iou=0.7)
    torch.cuda.synchronize()
    want, _ = N.non_max_suppression(y.cpu().numpy(), 0.01, 0.7, max_det=300)
    assert counts.tolist() == [len(w) for w in want] and rows[0, : len(want[0])].cpu().numpy().tobytes() == want[0].tobytes()
    assert _rel_l2(y.cpu(), y_ref) <= 2.0 * _rel_l2(O.forward(sd, x, "l", round_bf16=True, cfg="DBL2")[0], y_ref) + 3e-3


def test_sibling_graph_yolov13_area_attention():
    """Stock yolov13.yaml at n scale: strided / grouped stem, DSC3k2 with plain DSBottlenecks, A2C2f area attention
    (area 4 at P4, area 1 at P5), nn.Upsample -- teacher-forced per layer and free-running against the oracle (pinned to the reference's
    own yolov13n output, tests/test_oracle_golden.py), then the production engine's detections bit-exact against the oracle's NMS."""
    spec = resolve("yolov13n.yaml", "n", 3)
    sd = synth_state_dict(param_shapes(spec), 13)
    x = synth_images(2, 128, 160, 14)
    eng = DBLEngine(spec, sd, 2, 128, 160, use_graph=False, keep_layers=True)
    y_ref, feats_ref, outs_ref = O.forward(sd, x, "n", collect=True, cfg="V13")
    _, _, outs_b16 = O.forward(sd, x, "n", collect=True, round_bf16=True, cfg="V13")
    head_ref = torch.cat([f.reshape(2, f.shape[1], -1) for f in feats_ref], 2)
    forced = eng.debug_layers_forced(x.cuda(), outs_ref)
    errs = [(_rel_l2(o, r), i, spec.layers[i].kind) for i, (o, r) in enumerate(zip(forced, outs_ref + [head_ref]))]
    print("[yolov13n 128x160] teacher-forced rel-L2:", " ".join(f"{i}:{e:.3g}" for e, i, _ in errs))
    for e, i, kind in errs:
        assert e <= 1.5e-2, f"teacher-forced layer {i} ({kind}) relative L2 error {e:.3g}"
    outs = eng.debug_layers(x.cuda())
    fr = [(_rel_l2(o, r), _rel_l2(q, r), i) for i, (o, r, q) in enumerate(zip(outs, outs_ref, outs_b16)) if o is not None]
    print("[yolov13n 128x160] free-running rel-L2 (ours/oracle-bf16):", " ".join(f"{i}:{e:.3g}/{q:.3g}" for e, q, i in fr))
    for e, q, i in fr:
        assert e <= 2.0 * q + 3e-3, f"free-running layer {i} ({spec.layers[i].kind}): {e:.3g} vs bf16 drift {q:.3g}"
    m = DBLBackend.synthetic("n", 3, seed=13, cfg="yolov13n.yaml", device="cuda:0")
    y, _ = m(x.cuda())
    rows, counts = m.detect(x.cuda(), conf=0.01, iou=0.7)
    torch.cuda.synchronize()
    want, _ = N.non_max_suppression(y.cpu().numpy(), 0.01, 0.7, max_det=300)
    assert counts.tolist() == [len(w) for w in want]
    for i, w in enumerate(want):
        assert rows[i, : len(w)].cpu().numpy().tobytes() == w.tobytes()
    # the l scale adds A2C2f's gamma-scaled residual and mlp_ratio 1.5 (tasks.py:1088-1092): one teacher-forced pass over every layer
    spec_l = resolve("yolov13l.yaml", "l", 3)
    sd_l = synth_state_dict(param_shapes(spec_l), 1)
    xl = synth_images(1, 64, 64, 2)
    e_l = DBLEngine(spec_l, sd_l, 1, 64, 64, use_graph=False, keep_layers=True)
    _, feats_l, outs_l = O.forward(sd_l, xl, "l", collect=True, cfg="V13")
    head_l = torch.cat([f.reshape(1, f.shape[1], -1) for f in feats_l], 2)
    for i, (o, r) in enumerate(zip(e_l.debug_layers_forced(xl.cuda(), outs_l), outs_l + [head_l])):
        assert _rel_l2(o, r) <= 1.5e-2, f"yolov13l teacher-forced layer {i} ({spec_l.layers[i].kind}): {_rel_l2(o, r):.3g}"
