"""CPU-side checks: graph resolution vs the reference's parameter inventory, the C-ABI library loads and exports every
declared symbol, and the product fails loudly (no fallback) without a GPU."""
import ctypes
import json
import os
import re

import pytest
import torch

from yolo_dbl_b200 import _lib as L
from yolo_dbl_b200 import graph as G
from yolo_dbl_b200.synth import synth_images, synth_state_dict

from conftest import GOLDEN, ROOT


@pytest.mark.parametrize("scale,nc", [("n", 3), ("n", 80), ("s", 3), ("l", 3)])
def test_param_shapes_equal_reference_state_dict(scale, nc):
    ref = json.load(open(os.path.join(GOLDEN, f"shapes_{scale}_nc{nc}.json")))
    got = G.param_shapes(G.resolve(None, scale, nc))
    assert list(got) == list(ref)
    assert all(list(got[k]) == list(ref[k]) for k in ref)


def test_graph_channels_strides_match_survey_appendix_a():
    spec = G.resolve(None, "n", 80)
    assert [l.c2 for l in spec.layers[:35]] == [8, 16, 16, 32, 32, 64, 64, 128, 128, 128, 256, 256, 128, 128, 256, 128, 128, 256,
                                                256, 384, 128, 128, 128, 256, 64, 64, 64, 64, 192, 128, 128, 128, 384, 256, 256]
    assert spec.strides == [8, 16, 32] and spec.layers[12].he == 4 and G.resolve(None, "s", 3).layers[12].he == 8
    assert [l.extra["out_stride"] for l in spec.layers[12:15]] == [16, 8, 32]
    with pytest.raises(NotImplementedError):
        G.resolve({"scales": {"n": [1, 1, 1024]}, "backbone": [[-1, 1, "SPPF", [64]]], "head": []}, "n", 3)


def test_yaml_on_disk_is_accepted(tmp_path):
    import yaml

    p = tmp_path / "yolov13s_DBL.yaml"
    p.write_text(yaml.safe_dump(G.YOLO_DBL_CFG))
    assert G.scale_from_name(p.name) == "s"
    a, b = G.resolve(str(p), "s", 3), G.resolve(None, "s", 3)
    assert [l.c2 for l in a.layers] == [l.c2 for l in b.layers]


def test_synth_is_deterministic_and_calibrated():
    spec = G.resolve(None, "n", 3)
    a, b = synth_state_dict(G.param_shapes(spec), 0), synth_state_dict(G.param_shapes(spec), 0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert float(a["model.15.gate"]) == 0.5 and a["model.35.dfl.conv.weight"].flatten().tolist() == list(range(16))
    assert not torch.equal(a["model.0.bn.running_var"], synth_state_dict(G.param_shapes(spec), 0, calib=None)["model.0.bn.running_var"])
    x = synth_images(2, 64, 96, 3)
    assert x.shape == (2, 3, 64, 96) and 0 <= float(x.min()) and float(x.max()) <= 1 and torch.equal(x, synth_images(2, 64, 96, 3))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "dbl_sm100.h")).read()
    declared = set(re.findall(r"\b(dbl_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.dbl_version() >= 100
    assert ctypes.sizeof(L.Tensor) == 32 and ctypes.sizeof(L.Candidates) == 48


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    from yolo_dbl_b200 import DBLBackend, DblError, non_max_suppression

    m = DBLBackend.synthetic("n", 3)
    with pytest.raises(DblError):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(DblError):
        non_max_suppression(torch.zeros(1, 7, 100))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "yolo_dbl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "dbl_oracle" not in src and "nms_oracle" not in src and "import oracle" not in src, f


# ---- host-side mirrors of index arithmetic used inside the kernels (the GPU tests cover the kernels; these pin the maths) --------
def test_fastdiv_multiply_high_is_exact_within_its_guard():
    """dbl_common.cuh: q = umulhi(x, m), m = floor(2^32 / d) + 1 is exact while x * d < 2^32 (make_fastdiv leaves m = 0 otherwise)."""
    import random

    rng = random.Random(0)
    for d in [2, 3, 5, 6, 7, 11, 20, 54, 120, 255, 256, 1000, 4097, 65535]:
        m = (1 << 32) // d + 1
        lim = ((1 << 32) - 1) // d  # largest x with x * d < 2^32
        xs = [0, 1, d - 1, d, d + 1, lim - 1, lim] + [rng.randrange(0, lim + 1) for _ in range(2000)]
        for x in xs:
            if x < 0 or x * d >= (1 << 32):
                continue
            assert (x * m) >> 32 == x // d, (x, d)


def test_row_walk_staging_covers_every_chunk_once():
    """tc_ptx.cuh stage_tile16: threads keep a (column, group) and walk down the rows; every 16-byte chunk of the
    [rows][cols][G8] tile must be written exactly once, whatever the CTA size."""
    import math

    for rows, cols, g8, nt in [(10, 29, 2, 216), (14, 16, 2, 160), (38, 28, 2, 200), (18, 42, 1, 96), (14, 26, 8, 320), (10, 12, 16, 320),
                               (5, 70, 4, 128), (3, 300, 2, 256), (22, 22, 10, 416)]:
        row_items = cols * g8
        per_pass = min(row_items, nt)
        rpp = nt // per_pass
        seen = {}
        for tid in range(nt):
            r0 = math.floor((tid + 0.5) * (1.0 / per_pass))
            e_in = tid - r0 * per_pass
            if r0 >= rpp:
                continue
            e = e_in
            while e < row_items:
                for ly in range(r0, rows, rpp):
                    key = ly * row_items + e
                    seen[key] = seen.get(key, 0) + 1
                e += per_pass
        assert len(seen) == rows * row_items and set(seen.values()) == {1}, (rows, cols, g8, nt)
