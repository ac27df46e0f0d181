"""YOLO-DBL graph description: model yaml -> resolved layer table -> reference-compatible parameter shapes.

Mirrors what the reference's ``parse_model`` (``ultralytics/nn/tasks.py:947-1208``) decides for the module family
that ``cfg/models/v13/yolov13_DBL.yaml`` instantiates — depth/width scaling, ``make_divisible(.., 8)``, the HyperACE
hyperedge scaling, DownsampleConv / FullPAD / channel-preserving modules — without building any ``nn.Module``.
The B200 engine lowers this table to kernel launches; ``param_shapes`` lists the ``state_dict`` keys/shapes the
reference's ``DetectionModel`` would have, so weights move freely between the two.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Union

# cfg/models/v13/yolov13_DBL.yaml:1-53 as data (from, repeats, module, args); "l" is BASELINE.json's config 4, which
# injects YOLOv13-l width/depth into the same graph (the yaml itself defines n and s only).
YOLO_DBL_CFG = {
    "nc": 80,
    "scales": {"n": [0.50, 0.25, 114514], "s": [0.50, 0.50, 114514], "l": [1.00, 1.00, 114514]},
    "backbone": [
        [-1, 1, "Conv", [32, 3, 1]],
        [-1, 1, "Conv", [64, 3, 2]],
        [-1, 1, "Bottleneck", [64]],
        [-1, 1, "Conv", [128, 3, 2]],
        [-1, 2, "Bottleneck", [128]],
        [-1, 1, "Conv", [256, 3, 2]],
        [-1, 8, "Bottleneck", [256]],
        [-1, 1, "Conv", [512, 1, 1]],
        [-1, 1, "DSConv", [512, 3, 2]],
        [-1, 8, "DSC3k2", [512, True]],
        [-1, 1, "DSConv", [1024, 3, 2]],
        [-1, 4, "LSKblock", [1024]],
    ],
    "head": [
        [[7, 9, 11], 2, "HyperACE", [512, 8, True, True, 0.5, 1, "both", True]],
        [-1, 1, "DySample", []],
        [12, 1, "DownsampleConv", [True]],
        [[9, 12], 1, "FullPAD_Tunnel", []],
        [[7, 13], 1, "FullPAD_Tunnel", []],
        [[11, 14], 1, "FullPAD_Tunnel", []],
        [-1, 1, "DySample", []],
        [[-1, 15], 1, "Concat", [1]],
        [-1, 1, "Bottleneck", [512, True]],
        [[-1, 12], 1, "FullPAD_Tunnel", []],
        [-2, 1, "DySample", []],
        [[-1, 16], 1, "Concat", [1]],
        [-1, 1, "Bottleneck", [256, True]],
        [13, 1, "Conv", [256, 1, 1]],
        [[-2, -1], 1, "FullPAD_Tunnel", []],
        [-1, 1, "Conv", [256, 3, 2]],
        [[-1, 21], 1, "Concat", [1]],
        [-1, 1, "Bottleneck", [512, True]],
        [[-1, 12], 1, "FullPAD_Tunnel", []],
        [-2, 1, "Conv", [512, 3, 2]],
        [[-1, 17], 1, "Concat", [1]],
        [-1, 1, "DSC3k2", [1024, True]],
        [[-1, 14], 1, "FullPAD_Tunnel", []],
        [[26, 30, 34], 1, "Detect", ["nc"]],
    ],
}

# cfg/models/v13/yolov13_DBL2.yaml:1-52 -- the sibling graph of the same repository: layer 6 becomes a
# DSC3k2 stack, layers 9 and 33 become C3Ghost (GhostBottleneck / GhostConv, nn/modules/block.py:313-341, conv.py:184-197), the
# P4 / P3 neck blocks become DSC3k2; only the l and x scales are defined.
YOLO_DBL2_CFG = {
    "nc": 80,
    "scales": {"l": [1.00, 1.00, 114514], "x": [1.00, 1.50, 114514]},
    "backbone": [
        [-1, 1, "Conv", [32, 3, 1]],
        [-1, 1, "Conv", [64, 3, 2]],
        [-1, 1, "Bottleneck", [64]],
        [-1, 1, "Conv", [128, 3, 2]],
        [-1, 2, "Bottleneck", [128]],
        [-1, 1, "Conv", [256, 3, 2]],
        [-1, 8, "DSC3k2", [256, True]],
        [-1, 1, "Conv", [512, 1, 1]],
        [-1, 1, "DSConv", [512, 3, 2]],
        [-1, 8, "C3Ghost", [512, True]],
        [-1, 1, "DSConv", [1024, 3, 2]],
        [-1, 2, "LSKblock", [1024]],
    ],
    "head": [
        [[7, 9, 11], 2, "HyperACE", [512, 8, True, True, 0.5, 1, "both", True]],
        [-1, 1, "DySample", []],
        [12, 1, "DownsampleConv", [True]],
        [[9, 12], 1, "FullPAD_Tunnel", []],
        [[7, 13], 1, "FullPAD_Tunnel", []],
        [[11, 14], 1, "FullPAD_Tunnel", []],
        [-1, 1, "DySample", []],
        [[-1, 15], 1, "Concat", [1]],
        [-1, 1, "DSC3k2", [512, True]],
        [[-1, 12], 1, "FullPAD_Tunnel", []],
        [-2, 1, "DySample", []],
        [[-1, 16], 1, "Concat", [1]],
        [-1, 1, "Bottleneck", [256, True]],
        [13, 1, "DSC3k2", [256, 1, 1]],
        [[-2, -1], 1, "FullPAD_Tunnel", []],
        [-1, 1, "Conv", [256, 3, 2]],
        [[-1, 21], 1, "Concat", [1]],
        [-1, 1, "DSC3k2", [512, True]],
        [[-1, 12], 1, "FullPAD_Tunnel", []],
        [-2, 1, "Conv", [512, 3, 2]],
        [[-1, 17], 1, "Concat", [1]],
        [-1, 1, "C3Ghost", [1024, True]],
        [[-1, 14], 1, "FullPAD_Tunnel", []],
        [[26, 30, 34], 1, "Detect", ["nc"]],
    ],
}

# cfg/models/v13/yolov13.yaml:1-60 -- the stock YOLOv13 graph shipped beside the DBL variants: strided / grouped stem convolutions,
# DSC3k2 with plain DSBottlenecks (dsc3k = False, e = 0.25), A2C2f area-attention blocks (nn/modules/block.py:1170-1405), nearest
# upsampling instead of DySample, no LSKblock.
YOLOV13_CFG = {
    "nc": 80,
    "scales": {"n": [0.50, 0.25, 65536], "s": [0.50, 0.50, 65536], "l": [1.00, 1.00, 65536], "x": [1.00, 1.50, 65536]},
    "backbone": [
        [-1, 1, "Conv", [64, 3, 2]],
        [-1, 1, "Conv", [128, 3, 2, 1, 2]],
        [-1, 2, "DSC3k2", [256, False, 0.25]],
        [-1, 1, "Conv", [256, 3, 2, 1, 4]],
        [-1, 2, "DSC3k2", [512, False, 0.25]],
        [-1, 1, "DSConv", [512, 3, 2]],
        [-1, 4, "A2C2f", [512, True, 4]],
        [-1, 1, "DSConv", [1024, 3, 2]],
        [-1, 4, "A2C2f", [1024, True, 1]],
    ],
    "head": [
        [[4, 6, 8], 2, "HyperACE", [512, 8, True, True, 0.5, 1, "both"]],
        [-1, 1, "nn.Upsample", [None, 2, "nearest"]],
        [9, 1, "DownsampleConv", []],
        [[6, 9], 1, "FullPAD_Tunnel", []],
        [[4, 10], 1, "FullPAD_Tunnel", []],
        [[8, 11], 1, "FullPAD_Tunnel", []],
        [-1, 1, "nn.Upsample", [None, 2, "nearest"]],
        [[-1, 12], 1, "Concat", [1]],
        [-1, 2, "DSC3k2", [512, True]],
        [[-1, 9], 1, "FullPAD_Tunnel", []],
        [17, 1, "nn.Upsample", [None, 2, "nearest"]],
        [[-1, 13], 1, "Concat", [1]],
        [-1, 2, "DSC3k2", [256, True]],
        [10, 1, "Conv", [256, 1, 1]],
        [[21, 22], 1, "FullPAD_Tunnel", []],
        [-1, 1, "Conv", [256, 3, 2]],
        [[-1, 18], 1, "Concat", [1]],
        [-1, 2, "DSC3k2", [512, True]],
        [[-1, 9], 1, "FullPAD_Tunnel", []],
        [26, 1, "Conv", [512, 3, 2]],
        [[-1, 14], 1, "Concat", [1]],
        [-1, 2, "DSC3k2", [1024, True]],
        [[-1, 11], 1, "FullPAD_Tunnel", []],
        [[23, 27, 31], 1, "Detect", ["nc"]],
    ],
}

BUILTIN_CFGS = {"yolov13_DBL.yaml": YOLO_DBL_CFG, "yolov13_DBL2.yaml": YOLO_DBL2_CFG, "yolov13.yaml": YOLOV13_CFG}

SUPPORTED = {"Conv", "Bottleneck", "DSConv", "DSC3k2", "C3Ghost", "A2C2f", "LSKblock", "HyperACE", "DySample", "nn.Upsample",
             "DownsampleConv", "FullPAD_Tunnel", "Concat", "Detect"}


@dataclass
class LayerSpec:
    i: int
    src: List[int]          # absolute indices of the producing layers (-1 = the input image)
    kind: str
    n: int                  # repeats after depth scaling
    c1: Union[int, List[int]]
    c2: int
    k: int = 1
    s: int = 1
    shortcut: bool = True
    he: int = 0             # HyperACE: number of hyperedges
    extra: dict = field(default_factory=dict)


@dataclass
class GraphSpec:
    scale: str
    nc: int
    layers: List[LayerSpec]
    reg_max: int = 16

    @property
    def detect(self) -> LayerSpec:
        return self.layers[-1]

    @property
    def strides(self) -> List[int]:
        return self.layers[-1].extra["strides"]


def make_divisible(x: float, d: int = 8) -> int:
    return int(math.ceil(x / d) * d)


def load_cfg(cfg) -> dict:
    """Accepts None (built-in YOLO-DBL), a dict, the NAME of a built-in graph ('yolov13l_DBL2.yaml': the scale letter is ignored
    here) or a path to a yaml with the same schema."""
    if cfg is None:
        return YOLO_DBL_CFG
    if isinstance(cfg, dict):
        return cfg
    import os
    import re

    base = re.sub(r"(yolov\d+)[nslmx](_|\.)", r"\1\2", os.path.basename(str(cfg)))
    if base in BUILTIN_CFGS and not os.path.exists(str(cfg)):
        return BUILTIN_CFGS[base]
    import yaml

    with open(cfg, "r", errors="ignore") as f:
        return yaml.safe_load(f)


def builtin_cfg_name(name) -> str | None:
    """'yolov13l_DBL2.yaml' -> 'yolov13_DBL2.yaml' when that graph is built in, else None"""
    import os
    import re

    base = re.sub(r"(yolov\d+)[nslmx](_|\.)", r"\1\2", os.path.basename(str(name)))
    return base if base in BUILTIN_CFGS else None


def scale_from_name(name: str, default: str = "n") -> str:
    """'yolov13s_DBL.yaml' -> 's' (reference nn/tasks.py:1227-1242 guess_model_scale)."""
    import re

    m = re.search(r"yolov\d+([nslmx])", str(name))
    return m.group(1) if m else default


def resolve(cfg=None, scale: str = "n", nc: int | None = None) -> GraphSpec:
    d = load_cfg(cfg)
    nc = int(nc if nc is not None else d.get("nc", 80))
    scales = d["scales"]
    if scale not in scales:
        raise ValueError(f"scale {scale!r} not in {list(scales)}")
    depth, width, max_ch = scales[scale]
    ch: List[int] = []
    strides: List[int] = []   # cumulative stride of each layer's output
    layers: List[LayerSpec] = []
    rows = list(d["backbone"]) + list(d["head"])
    for i, (f, n, kind, args) in enumerate(rows):
        if kind not in SUPPORTED:
            raise NotImplementedError(f"layer {i}: module {kind} is outside the YOLO-DBL hot path")
        args = list(args)
        n = max(round(n * depth), 1) if n > 1 else n
        fl = [f] if isinstance(f, int) else list(f)
        src = [(i + j if j < 0 else j) for j in fl]
        if i == 0:
            src = [-1]
        cin = [3 if j < 0 else ch[j] for j in src]
        sin = [1 if j < 0 else strides[j] for j in src]
        L = LayerSpec(i, src, kind, n, cin[0], cin[0])
        st = sin[0]
        if kind in ("Conv", "DSConv"):
            L.c2 = make_divisible(min(args[0], max_ch) * width)
            L.k, L.s = int(args[1]), int(args[2])
            L.extra["g"] = int(args[4]) if (kind == "Conv" and len(args) > 4) else 1  # Conv(c1, c2, k, s, p, g) (conv.py:39-52)
            if len(args) > 3 and args[3] is not None and int(args[3]) != L.k // 2:
                raise NotImplementedError(f"layer {i}: padding {args[3]} != k//2")
            st *= L.s
        elif kind == "Bottleneck":
            L.c2 = make_divisible(min(args[0], max_ch) * width)
            L.shortcut = bool(args[1]) if len(args) > 1 else True
        elif kind == "DSC3k2":
            L.c2 = make_divisible(min(args[0], max_ch) * width)
            L.extra["dsc3k"] = bool(args[1]) if len(args) > 1 else False
            L.extra["e"] = float(args[2]) if len(args) > 2 else 0.5  # DSC3k2(c1, c2, n, dsc3k, e=0.5, ...) (block.py:1505-1580)
        elif kind == "A2C2f":  # A2C2f(c1, c2, n, a2, area, [residual, mlp_ratio for l/x]) (tasks.py:1088-1092, block.py:1350-1405)
            L.c2 = make_divisible(min(args[0], max_ch) * width)
            L.extra.update(a2=bool(args[1]), area=int(args[2]), residual=scale in "lx", mlp_ratio=1.5 if scale in "lx" else 2.0)
            if not L.extra["a2"]:
                raise NotImplementedError("A2C2f(a2=False) is not used by the v13 graphs")
            if int(L.c2 * 0.5) % 32:
                raise ValueError("Dimension of ABlock must be a multiple of 32")
        elif kind == "nn.Upsample":
            if list(args) != [None, 2, "nearest"]:
                raise NotImplementedError("only nn.Upsample(None, 2, 'nearest')")
            st = max(st // 2, 1)
        elif kind == "C3Ghost":
            L.c2 = make_divisible(min(args[0], max_ch) * width)
            L.shortcut = bool(args[1]) if len(args) > 1 else True  # unused by GhostBottleneck (its own identity shortcut)
        elif kind == "HyperACE":
            L.c1 = cin[1]
            L.c2 = make_divisible(min(args[0], max_ch) * width)
            he = args[1]
            if scale in "n":        # reference: `if scale in "n"` / `elif scale in "x"` (tasks.py:1126-1130)
                he = int(he * 0.5)
            elif scale in "x":
                he = int(he * 1.5)
            L.he = he
            L.extra.update(dsc3k=bool(args[2]), shortcut=bool(args[3]), e1=float(args[4]), e2=float(args[5]),
                           context=args[6], channel_adjust=bool(args[7]) if len(args) > 7 else True, cins=cin)  # default: block.py:1874
            if not (L.extra["dsc3k"] and L.extra["context"] == "both" and L.extra["channel_adjust"]):
                raise NotImplementedError("HyperACE variant outside YOLO-DBL")
            if not L.extra["shortcut"]:  # the lowering always emits the residual DSBottleneck (DSBottleneck.add = shortcut and c1 == c2)
                raise NotImplementedError("HyperACE(shortcut=False) is not used by YOLO-DBL")
            if sum(cin) != 4 * L.c1:
                raise ValueError("FuseModule(channel_adjust=True) needs c(P3)+c(P4)+c(P5) == 4*c(P4)")
            st = sin[1]
            L.n, n = n, 1
        elif kind == "DownsampleConv":
            L.c2 = 2 * cin[0]
            st *= 2
        elif kind == "DySample":
            st = max(st // 2, 1)
        elif kind == "Concat":
            L.c1, L.c2 = cin, sum(cin)
        elif kind == "FullPAD_Tunnel":
            L.c1 = cin
            if cin[0] != cin[1]:
                raise ValueError(f"layer {i}: FullPAD_Tunnel inputs differ in channels {cin}")
        elif kind == "Detect":
            L.c1, L.c2 = cin, 0
            L.extra["strides"] = sin
        L.extra["out_stride"] = st
        layers.append(L)
        ch.append(L.c2)
        strides.append(st)
    return GraphSpec(scale, nc, layers)


# ---------------------------------------------------------------------------------------------------------------------
# reference-compatible parameter inventory
# ---------------------------------------------------------------------------------------------------------------------
def _conv(sh: Dict[str, list], p: str, c1: int, c2: int, k: int, g: int = 1):
    sh[p + ".conv.weight"] = [c2, c1 // g, k, k]
    for nm in ("weight", "bias", "running_mean", "running_var"):
        sh[f"{p}.bn.{nm}"] = [c2]
    sh[p + ".bn.num_batches_tracked"] = []


def _dsconv(sh, p, c1, c2, k):
    sh[p + ".dw.weight"] = [c1, 1, k, k]
    sh[p + ".pw.weight"] = [c2, c1, 1, 1]
    for nm in ("weight", "bias", "running_mean", "running_var"):
        sh[f"{p}.bn.{nm}"] = [c2]
    sh[p + ".bn.num_batches_tracked"] = []


def _dsc3k(sh, p, c1, c2, n, e, k1=3, k2=7):
    c_ = int(c2 * e)
    _conv(sh, p + ".cv1", c1, c_, 1)
    _conv(sh, p + ".cv2", c1, c_, 1)
    _conv(sh, p + ".cv3", 2 * c_, c2, 1)
    for j in range(n):
        _dsconv(sh, f"{p}.m.{j}.cv1", c_, c_, k1)
        _dsconv(sh, f"{p}.m.{j}.cv2", c_, c_, k2)


def _c3ah(sh, p, c, he):
    _conv(sh, p + ".cv1", c, c, 1)
    _conv(sh, p + ".cv2", c, c, 1)
    g = p + ".m.hgnn"
    sh[g + ".edge_generator.prototype_base"] = [he, c]
    sh[g + ".edge_generator.context_net.weight"] = [he * c, 2 * c]
    sh[g + ".edge_generator.context_net.bias"] = [he * c]
    sh[g + ".edge_generator.pre_head_proj.weight"] = [c, c]
    sh[g + ".edge_generator.pre_head_proj.bias"] = [c]
    sh[g + ".edge_proj.0.weight"] = [c, c]
    sh[g + ".edge_proj.0.bias"] = [c]
    sh[g + ".node_proj.0.weight"] = [c, c]
    sh[g + ".node_proj.0.bias"] = [c]
    _conv(sh, p + ".cv3", 2 * c, c, 1)


def detect_channels(nc: int, ch: Sequence[int]):
    """Detect.__init__ (nn/modules/head.py:86): c2 = box-branch width, c3 = class-branch width."""
    return max(16, ch[0] // 4, 64), max(ch[0], min(nc, 100))


def param_shapes(spec: GraphSpec) -> Dict[str, list]:
    """state_dict key -> shape, in the reference's registration order."""
    sh: Dict[str, list] = {}
    for L in spec.layers:
        p = f"model.{L.i}"
        rep = (lambda r, L=L, p=p: f"{p}.{r}" if L.n > 1 else p)
        if L.kind == "Conv":
            _conv(sh, p, L.c1, L.c2, L.k, g=L.extra.get("g", 1))
        elif L.kind == "DSConv":
            _dsconv(sh, p, L.c1, L.c2, L.k)
        elif L.kind == "Bottleneck":
            for r in range(L.n):
                c1 = L.c1 if r == 0 else L.c2
                _conv(sh, rep(r) + ".cv1", c1, int(L.c2 * 0.5), 3)
                _conv(sh, rep(r) + ".cv2", int(L.c2 * 0.5), L.c2, 3)
        elif L.kind == "DSC3k2":
            c = int(L.c2 * L.extra.get("e", 0.5))
            _conv(sh, p + ".cv1", L.c1, 2 * c, 1)
            _conv(sh, p + ".cv2", (2 + L.n) * c, L.c2, 1)
            for j in range(L.n):
                if L.extra.get("dsc3k", True):
                    _dsc3k(sh, f"{p}.m.{j}", c, c, 2, 1.0)
                else:  # plain DSBottleneck(c, c, e=1.0, k1=3, k2=7) (block.py:1565-1577)
                    _dsconv(sh, f"{p}.m.{j}.cv1", c, c, 3)
                    _dsconv(sh, f"{p}.m.{j}.cv2", c, c, 7)
        elif L.kind == "A2C2f":
            c_ = int(L.c2 * 0.5)
            hid = int(c_ * L.extra["mlp_ratio"])
            if L.extra["residual"]:
                sh[p + ".gamma"] = [L.c2]  # a parameter of the module itself: precedes the children in state_dict order
            _conv(sh, p + ".cv1", L.c1, c_, 1)
            _conv(sh, p + ".cv2", (1 + L.n) * c_, L.c2, 1)
            for j in range(L.n):
                for b in range(2):  # nn.Sequential(ABlock, ABlock)
                    q = f"{p}.m.{j}.{b}"
                    _conv(sh, q + ".attn.qkv", c_, 3 * c_, 1)
                    _conv(sh, q + ".attn.proj", c_, c_, 1)
                    _conv(sh, q + ".attn.pe", c_, c_, 7, g=c_)
                    _conv(sh, q + ".mlp.0", c_, hid, 1)
                    _conv(sh, q + ".mlp.1", hid, c_, 1)
        elif L.kind == "C3Ghost":  # C3 (block.py:259-273) whose m = n GhostBottleneck(c_, c_) (block.py:313-341)
            c_ = int(L.c2 * 0.5)
            _conv(sh, p + ".cv1", L.c1, c_, 1)
            _conv(sh, p + ".cv2", L.c1, c_, 1)
            _conv(sh, p + ".cv3", 2 * c_, L.c2, 1)
            for j in range(L.n):
                q = f"{p}.m.{j}.conv"
                h = c_ // 2                      # GhostBottleneck hidden width
                _conv(sh, q + ".0.cv1", c_, h // 2, 1)             # GhostConv(c_, h): primary 1x1 ...
                _conv(sh, q + ".0.cv2", h // 2, h // 2, 5, g=h // 2)  # ... and cheap 5x5 depthwise (conv.py:184-197)
                _conv(sh, q + ".2.cv1", h, c_ // 2, 1)              # GhostConv(h, c_, act=False)
                _conv(sh, q + ".2.cv2", c_ // 2, c_ // 2, 5, g=c_ // 2)
        elif L.kind == "LSKblock":
            d = L.c1
            for r in range(L.n):
                q = rep(r)
                sh[q + ".conv0.weight"], sh[q + ".conv0.bias"] = [d, 1, 5, 5], [d]
                sh[q + ".conv_spatial.weight"], sh[q + ".conv_spatial.bias"] = [d, 1, 7, 7], [d]
                sh[q + ".conv1.weight"], sh[q + ".conv1.bias"] = [d // 2, d, 1, 1], [d // 2]
                sh[q + ".conv2.weight"], sh[q + ".conv2.bias"] = [d // 2, d, 1, 1], [d // 2]
                sh[q + ".conv_squeeze.weight"], sh[q + ".conv_squeeze.bias"] = [2, 2, 7, 7], [2]
                sh[q + ".conv.weight"], sh[q + ".conv.bias"] = [d, d // 2, 1, 1], [d]
        elif L.kind == "HyperACE":
            c = int(L.c2 * L.extra["e1"])
            _conv(sh, p + ".cv1", L.c1, 3 * c, 1)
            _conv(sh, p + ".cv2", (4 + L.n) * c, L.c2, 1)
            for j in range(L.n):
                _dsc3k(sh, f"{p}.m.{j}", c, c, 2, 0.5)
            _conv(sh, p + ".fuse.conv_out", 4 * L.c1, L.c1, 1)
            _c3ah(sh, p + ".branch1", c, L.he)
            _c3ah(sh, p + ".branch2", c, L.he)
        elif L.kind == "DySample":
            sh[p + ".init_pos"] = [1, 32, 1, 1]
            sh[p + ".offset.weight"], sh[p + ".offset.bias"] = [32, L.c1, 1, 1], [32]
        elif L.kind == "nn.Upsample":
            pass
        elif L.kind == "DownsampleConv":
            _conv(sh, p + ".channel_adjust", L.c1, L.c2, 1)
        elif L.kind == "FullPAD_Tunnel":
            sh[p + ".gate"] = []
        elif L.kind == "Detect":
            c2, c3 = detect_channels(spec.nc, L.c1)
            for i, x in enumerate(L.c1):
                _conv(sh, f"{p}.cv2.{i}.0", x, c2, 3)
                _conv(sh, f"{p}.cv2.{i}.1", c2, c2, 3)
                sh[f"{p}.cv2.{i}.2.weight"], sh[f"{p}.cv2.{i}.2.bias"] = [4 * spec.reg_max, c2, 1, 1], [4 * spec.reg_max]
            for i, x in enumerate(L.c1):
                _conv(sh, f"{p}.cv3.{i}.0.0", x, x, 3, g=x)
                _conv(sh, f"{p}.cv3.{i}.0.1", x, c3, 1)
                _conv(sh, f"{p}.cv3.{i}.1.0", c3, c3, 3, g=c3)
                _conv(sh, f"{p}.cv3.{i}.1.1", c3, c3, 1)
                sh[f"{p}.cv3.{i}.2.weight"], sh[f"{p}.cv3.{i}.2.bias"] = [spec.nc, c3, 1, 1], [spec.nc]
            sh[p + ".dfl.conv.weight"] = [1, spec.reg_max, 1, 1]
    return sh
