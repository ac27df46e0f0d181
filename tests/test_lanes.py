"""CPU test of the engine's dependency planner (engine._plan_lanes): every hazard between two launches of the real YOLO-DBL
launch list must be ordered by the lane plan (same-lane program order or a chain of cross-lane event waits).

The engine is lowered against a stand-in for the CUDA library (no kernel runs; buffers live on the CPU), which is enough for
the planner: it only looks at the address ranges of the dbl_tensor arguments."""
import types

import pytest
import torch

from yolo_dbl_b200 import _lib as L
from yolo_dbl_b200 import engine as E
from yolo_dbl_b200 import param_shapes, resolve, synth_state_dict


class _Fn:
    def __init__(self, name):
        self.__name__ = name

    def __call__(self, *a):
        return 0


class _FakeLib:
    def __getattr__(self, n):
        if n == "dbl_mma_packed_bytes":
            return lambda co, ci, t: ((t + 1) // 2 if ci == 8 else t * (ci // 16)) * (co // 8) * 256
        if n == "dbl_hypergraph_workspace_bytes":
            return lambda *a: 1024
        if n in ("dbl_head_decode_supported", "dbl_dsconv_supported"):
            return lambda *a: 1
        if n == "dbl_conv2d_which":
            return lambda *a: 1
        f = _Fn(n)
        setattr(self, n, f)
        return f


@pytest.fixture
def fake_engine(monkeypatch):
    lib = _FakeLib()
    monkeypatch.setattr(L, "require_device", lambda: lib)
    monkeypatch.setattr(L, "current_stream", lambda: 0)
    monkeypatch.setattr(L, "check", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

    class _Dev:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    monkeypatch.setattr(torch.cuda, "device", lambda d: _Dev())
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: None)

    def build(scale="n", nc=80, lanes=8, **kw):
        spec = resolve(None, scale, nc)
        sd = synth_state_dict(param_shapes(spec), 0)
        return E.DBLEngine(spec, sd, 2, 64, 64, device="cpu", use_graph=True, lanes=lanes, **kw)

    return build


def _ordered(plan):
    """reach[j] = bitset of launches guaranteed to complete before launch j starts under the plan"""
    last_on_lane, reach = {}, []
    for j, (lane, waits) in enumerate(plan):
        r = 0
        preds = list(waits) + ([last_on_lane[lane]] if lane in last_on_lane else [])
        for i in preds:
            r |= reach[i] | (1 << i)
        reach.append(r)
        last_on_lane[lane] = j
    return reach


@pytest.mark.parametrize("scale,nc,lanes", [("n", 80, 8), ("n", 3, 3), ("s", 3, 8), ("l", 3, 2)])
def test_every_hazard_is_ordered_by_the_lane_plan(fake_engine, scale, nc, lanes):
    eng = fake_engine(scale, nc, lanes)
    n = len(eng.ops)
    assert n > 100 and len(eng.plan) == n and len(eng.op_rw) == n
    assert len({lane for lane, _ in eng.plan}) > 1, "the graph has independent branches: more than one lane must be used"
    reach = _ordered(eng.plan)
    ov = E.DBLEngine._overlap
    hazards = 0
    for j in range(n):
        rj, wj = eng.op_rw[j]
        for i in range(j):
            ri, wi = eng.op_rw[i]
            if any(ov(w, r) for w in wi for r in rj) or any(ov(w, w2) for w in wi for w2 in wj) or any(ov(r, w) for r in ri for w in wj):
                hazards += 1
                assert (reach[j] >> i) & 1, f"launch {j} ({eng.ops[j][2]}) may run before / with launch {i} ({eng.ops[i][2]})"
    assert hazards >= n - 1  # at least a spanning chain of true dependencies exists


def test_overlap_of_channel_slices():
    ov = E.DBLEngine._overlap
    base, stride = 1 << 20, 256  # one buffer, 128 bf16 channels per pixel, 100 pixels
    def view(off, c):
        return (base + off * 2, base + (99 * 128 + off + c) * 2, stride, c * 2)
    assert not ov(view(0, 64), view(64, 64))  # disjoint channel slices of the same pixels
    assert ov(view(0, 64), view(32, 64)) and ov(view(0, 128), view(64, 8)) and ov(view(8, 8), view(8, 8))
    assert not ov(view(0, 64), (base + 10 ** 6, base + 10 ** 6 + 500, stride, 128))  # different allocation
    assert ov(view(0, 64), (base, base + 99 * 64 * 2, 128, 128))  # same memory seen with another pitch: conservative
