"""Pin the CPU oracle against outputs of the REAL reference (tests/golden/, made by oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

import dbl_oracle as O
import nms_oracle as N
from yolo_dbl_b200.graph import param_shapes, resolve
from yolo_dbl_b200.synth import synth_images, synth_state_dict

from conftest import GOLDEN


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def _case(name):
    g = _load(f"fwd_{name}.npz")
    meta = json.loads(str(g["meta"]))
    shapes = json.load(open(os.path.join(GOLDEN, f"shapes_{meta['scale']}_nc{meta['nc']}.json")))
    sd = synth_state_dict(shapes, meta["wseed"])
    x = synth_images(meta["b"], meta["h"], meta["w"], meta["xseed"])
    return g, meta, sd, x


@pytest.mark.parametrize("name", ["n_nc3_small", "n_nc80_small", "s_nc3_small", "l_nc3_small"])
def test_forward_matches_reference_small(name):
    g, meta, sd, x = _case(name)
    y, feats, outs = O.forward(sd, x, meta["scale"], collect=True)
    ref = torch.from_numpy(g["y_unfused"])  # the oracle applies BN explicitly, like the unfused reference
    assert y.shape == ref.shape
    torch.testing.assert_close(y, ref, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(y, torch.from_numpy(g["y"]), rtol=1e-3, atol=1e-3)  # reference after fuse()
    for li, f in enumerate(feats):
        samp = f.reshape(-1)[torch.from_numpy(g[f"feat{li}_index"])]
        torch.testing.assert_close(samp, torch.from_numpy(g[f"feat{li}_samples"]), rtol=1e-3, atol=1e-3)
    if "layer0_stats" in g:
        for i, o in enumerate(outs):
            assert tuple(o.shape) == tuple(g[f"layer{i}_shape"]), f"layer {i}"
            samp = o.reshape(-1)[torch.from_numpy(g[f"layer{i}_index"])]
            torch.testing.assert_close(samp, torch.from_numpy(g[f"layer{i}_samples"]), rtol=2e-4, atol=2e-4,
                                       msg=lambda m, i=i: f"layer {i}: {m}")
            st = g[f"layer{i}_stats"]
            assert abs(float(o.double().mean()) - st[0]) < 1e-4 + 1e-4 * abs(st[0])
            assert abs(float(o.double().std()) - st[1]) < 1e-4 + 1e-4 * abs(st[1])


def test_forward_matches_reference_640():
    g, meta, sd, x = _case("n_nc3_640")
    y, feats, outs = O.forward(sd, x, "n", collect=True)
    ys = y[:, :, :: meta["y_stride"]]
    torch.testing.assert_close(ys, torch.from_numpy(g["y_unfused"]), rtol=2e-4, atol=5e-3)
    for i, o in enumerate(outs):
        samp = o.reshape(-1)[torch.from_numpy(g[f"layer{i}_index"])]
        torch.testing.assert_close(samp, torch.from_numpy(g[f"layer{i}_samples"]), rtol=5e-4, atol=5e-4,
                                   msg=lambda m, i=i: f"layer {i}: {m}")
    # NMS oracle on the reference's own decoded y == the reference's NMS rows, bit for bit
    for conf in (0.25, 0.001):
        ref = _load(f"nms_model640_conf{conf}.npz")
        rows, idx = N.non_max_suppression(g["y"], conf, 0.7, max_det=300)
        assert rows[0].tobytes() == ref["rows"].astype(np.float32).tobytes()
        assert len(set(idx[0].tolist())) == len(idx[0])


def test_layer_table_matches_reference_docstring_shapes():
    # shape-level "known answers" of the reference (block.py docstrings) + Appendix A channel table
    L = O.build_layers("n")
    assert [l.c2 for l in L[:15]] == [8, 16, 16, 32, 32, 64, 64, 128, 128, 128, 256, 256, 128, 128, 256]
    assert [l.n for l in L if l.kind in ("Bottleneck", "DSC3k2", "LSKblock", "HyperACE")][:6] == [1, 1, 4, 4, 2, 1]
    assert L[12].extra["he"] == 4 and O.build_layers("s")[12].extra["he"] == 8 and O.build_layers("l")[12].extra["he"] == 8
    assert L[35].c1 == [64, 128, 256]


def test_init_pos_matches_reference_buffer():
    from yolo_dbl_b200.synth import _init_pos

    torch.testing.assert_close(_init_pos(), O.dysample_init_pos())


# ------------------------------------------------------------------------------------------------ NMS
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
def test_nms_oracle_bit_exact_vs_reference(name):
    g = _load(f"nms_{name}.npz")
    meta = json.loads(str(g["meta"]))
    pred = _synth_prediction(**meta["pred_args"]).numpy()
    rows, idx = N.non_max_suppression(pred, **meta["kw"])
    assert [r.shape[0] for r in rows] == g["counts"].tolist()
    got = np.concatenate(rows, 0) if rows else np.zeros((0, 6), np.float32)
    assert got.tobytes() == g["rows"].astype(np.float32).tobytes(), "rows differ from the reference bit pattern"
    for r, i in zip(rows, idx):
        assert len(i) == r.shape[0]


def test_nms_kernel_matches_torchvision_random_and_ties():
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(1, 400))
        xy = rng.uniform(0, 200, (n, 2)).astype(np.float32)
        wh = rng.uniform(1, 80, (n, 2)).astype(np.float32)
        if trial % 3 == 0:
            xy, wh = np.round(xy), np.round(wh)  # integer boxes: exact-threshold IoUs
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        scores = rng.uniform(0, 1, n).astype(np.float32)
        if trial % 2 == 0:
            scores = np.round(scores * 8) / 8  # many ties
        thr = float(rng.choice([0.3, 0.45, 0.5, 0.7]))
        want = tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
        np.testing.assert_array_equal(N.nms(boxes, scores, thr), want)


# ---- letterbox / scale_boxes (image-source pre/post-process): oracle vs the reference's LetterBox (cv2) and ops.scale_boxes
LETTERBOX_CASES = ["down_wide", "up_tall", "same", "odd", "square_down", "up2x", "down2x", "up_both", "thin", "auto_wide", "auto_tall"]


@pytest.mark.parametrize("name", LETTERBOX_CASES)
def test_letterbox_oracle_matches_reference_cv2(name):
    import letterbox_oracle as LB

    g = _load(f"letterbox_{name}.npz")
    shape = tuple(int(v) for v in g["shape"])
    auto = name.startswith("auto_")  # minimum-rectangle mode of BasePredictor.pre_transform (same-shape batches)
    got = LB.letterbox(g["img"], shape, auto=auto)
    assert got.shape == g["out"].shape
    # integer work: byte for byte what the reference's LetterBox (cv2.resize INTER_LINEAR + copyMakeBorder) produced
    np.testing.assert_array_equal(got, g["out"])
    # padding is exact
    new_h, new_w, top, left = LB.letterbox_params(g["img"].shape[0], g["img"].shape[1], *shape, auto=auto)[:4]
    pad = np.ones(got.shape[:2], bool)
    pad[top:top + new_h, left:left + new_w] = False
    assert (got[pad] == 114).all() and (g["out"][pad] == 114).all()
    # scale_boxes + clip_boxes: fp32 arithmetic, bit for bit what the reference's ops.scale_boxes returned
    np.testing.assert_array_equal(LB.scale_boxes(got.shape[:2], g["boxes"], g["img"].shape[:2]), g["scaled"])


# ---- round 2: validator path (a-priori labels, match_predictions) and the benchmarked nc=80 instantiation ----------------
@pytest.mark.parametrize("name", ["single", "multi"])
def test_nms_oracle_labels_branch_matches_reference(name):
    """ops.non_max_suppression(labels=lb) (utils/ops.py:255-261): rows identical to the reference's output, byte for byte"""
    g = _load(f"nms_labels_{name}.npz")
    lc = g["label_counts"].tolist()
    labels, o = [], 0
    for n in lc:
        labels.append(g["labels"][o:o + n])
        o += n
    rows, idx = N.non_max_suppression(g["pred"], 0.4, 0.6, multi_label=bool(g["multi_label"]), labels=labels, max_det=300)
    assert [len(r) for r in rows] == g["counts"].tolist()
    assert np.concatenate(rows).tobytes() == g["rows"].tobytes()
    # the case is only meaningful if label rows (score exactly 1, index >= A) survive
    A = g["pred"].shape[2]
    assert any((i >= A).any() for i in idx) and all(float(r[:, 4].max()) == 1.0 for r, n in zip(rows, lc) if n)


@pytest.mark.parametrize("name", ["typical", "crowded", "empty", "one"])
def test_match_oracle_matches_reference(name):
    """box_iou + DetectionValidator.match_predictions of the reference vs the numpy restatement"""
    import match_oracle as M

    g = _load(f"match_{name}.npz")
    if len(g["gt"]):
        iou = M.box_iou(g["gt"], g["det"][:, :4])
        assert iou.tobytes() == g["iou"].tobytes()  # strict fp32, same operation order
    got = M.process_batch(g["det"], g["gt"], g["gt_cls"], g["iouv"])
    np.testing.assert_array_equal(got, g["correct"])


def test_oracle_forward_nc80_640_matches_reference_on_the_benchmarked_weights():
    """The instantiation bench.py measures (n, nc=80, recipe seed 0 + calibration table, bench images): oracle vs the REAL
    reference's fused forward and NMS."""
    g = _load("fwd_n_nc80_640.npz")
    spec = resolve(None, "n", 80)
    sd = synth_state_dict(param_shapes(spec), 0)
    x = synth_images(8, 640, 640, seed=100)[2:4]
    y, _ = O.forward(sd, x, "n")
    # the oracle applies BatchNorm explicitly like the UNFUSED reference: same bound as test_forward_matches_reference_640
    torch.testing.assert_close(y[0, :, ::4], torch.from_numpy(g["ya_unfused_s4"]), rtol=2e-4, atol=5e-3)
    # against the fused model (what predict runs) the fp32 re-association of fuse() shows: <= 0.25 px / 5e-3 in score
    d = (y[0] - torch.from_numpy(g["ya"])).abs()
    assert float(d[:4].max()) <= 0.25 and float(d[4:].max()) <= 5e-3, (float(d[:4].max()), float(d[4:].max()))
    d1 = (y[1, :4] - torch.from_numpy(g["yb_box"])).abs()
    best, cls = y[1, 4:].max(0)
    assert float(d1.max()) <= 0.25 and float((best - torch.from_numpy(g["yb_best"])).abs().max()) <= 5e-3
    rows, _ = N.non_max_suppression(np.stack([g["ya"], y[1].numpy()]), 0.25, 0.7, max_det=300)
    c = g["counts"].tolist()
    assert len(rows[0]) == c[0] and rows[0].tobytes() == g["rows"][: c[0]].tobytes()  # image 0: the reference's own y -> identical rows
    assert abs(len(rows[1]) - c[1]) <= 2  # image 1: the oracle's y (fp32 noise may flip a box at the threshold)


def test_dbl2_graph_matches_reference_inventory_and_forward():
    """Sibling graph yolov13l_DBL2.yaml: parameter inventory == the reference's state_dict, oracle forward (C3Ghost,
    GhostBottleneck, GhostConv, DSC3k2 with e = 1) == the reference's output and per-layer samples."""
    shapes = json.load(open(os.path.join(GOLDEN, "shapes_l_DBL2_nc3.json")))
    spec = resolve("yolov13l_DBL2.yaml", "l", 3)
    mine = param_shapes(spec)
    assert list(mine.keys()) == list(shapes.keys()) and all(list(mine[k]) == list(shapes[k]) for k in shapes)
    assert [L.kind for L in spec.layers if L.kind in ("C3Ghost", "DSC3k2")] == ["DSC3k2", "C3Ghost", "DSC3k2", "DSC3k2", "DSC3k2", "C3Ghost"]
    g = _load("fwd_l_DBL2_nc3_small.npz")
    sd = synth_state_dict(shapes, 9)
    x = synth_images(1, 64, 96, seed=10)
    y, feats, outs = O.forward(sd, x, "l", collect=True, cfg="DBL2")
    torch.testing.assert_close(y, torch.from_numpy(g["y_unfused"]), rtol=2e-4, atol=2e-3)
    torch.testing.assert_close(y, torch.from_numpy(g["y"]), rtol=1e-3, atol=2e-2)  # reference after fuse()
    for i, o in enumerate(outs):
        assert tuple(o.shape) == tuple(g[f"layer{i}_shape"]), f"layer {i}"
        samp = o.reshape(-1)[torch.from_numpy(g[f"layer{i}_index"])]
        torch.testing.assert_close(samp, torch.from_numpy(g[f"layer{i}_samples"]), rtol=5e-4, atol=5e-4, msg=lambda m, i=i: f"layer {i}: {m}")


def test_yolov13_graph_matches_reference_inventory_and_forward():
    """Stock yolov13.yaml: parameter inventory == the reference's state_dict; oracle forward (grouped convs, DSC3k2 with
    plain DSBottlenecks, A2C2f / ABlock / AAttn with area 4 and 1, nn.Upsample) == the reference's output and per-layer samples."""
    shapes = json.load(open(os.path.join(GOLDEN, "shapes_n_V13_nc3.json")))
    spec = resolve("yolov13n.yaml", "n", 3)
    mine = param_shapes(spec)
    assert list(mine.keys()) == list(shapes.keys()) and all(list(mine[k]) == list(shapes[k]) for k in shapes)
    assert [L.kind for L in spec.layers[:9]] == ["Conv", "Conv", "DSC3k2", "Conv", "DSC3k2", "DSConv", "A2C2f", "DSConv", "A2C2f"]
    g = _load("fwd_n_V13_nc3_small.npz")
    sd = synth_state_dict(shapes, 13)
    x = synth_images(1, 128, 160, seed=14)
    y, feats, outs = O.forward(sd, x, "n", collect=True, cfg="V13")
    for i, o in enumerate(outs):
        assert tuple(o.shape) == tuple(g[f"layer{i}_shape"]), f"layer {i}"
        samp = o.reshape(-1)[torch.from_numpy(g[f"layer{i}_index"])]
        torch.testing.assert_close(samp, torch.from_numpy(g[f"layer{i}_samples"]), rtol=5e-4, atol=5e-4, msg=lambda m, i=i: f"layer {i}: {m}")
    torch.testing.assert_close(y, torch.from_numpy(g["y_unfused"]), rtol=2e-4, atol=5e-3)
    # the l / x variants add the gamma residual and mlp_ratio 1.5 (tasks.py:1088-1092): shapes only (the model is 100+ MB)
    big = param_shapes(resolve("yolov13l.yaml", "l", 3))
    assert big["model.6.gamma"] == [512] and big["model.6.m.0.0.mlp.0.conv.weight"][0] == int(256 * 1.5)
