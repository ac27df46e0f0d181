"""yolo_dbl_b200 — B200-native (sm_100a) inference path for YOLO-DBL behind the reference's ultralytics-style API.

    from yolo_dbl_b200 import YOLO, DBLBackend, non_max_suppression

Host code is Python/PyTorch (buffers, streams, CUDA graphs); every arithmetic step is a hand-written CUDA kernel in
``lib/libdbl_sm100.so`` reached through the C ABI in ``include/dbl_sm100.h``.  No CPU fallback exists.
"""
from ._lib import DblError, LIB_PATH  # noqa: F401
from .graph import GraphSpec, param_shapes, resolve  # noqa: F401
from .synth import synth_images, synth_state_dict  # noqa: F401


def __getattr__(name):  # backend / engine import torch.nn and touch the library lazily
    if name in ("DBLBackend", "YOLO", "Results", "Boxes", "non_max_suppression", "nms_raw"):
        from . import backend

        return getattr(backend, name)
    if name in ("DBLValidator", "match_predictions"):
        from . import validator

        return getattr(validator, name)
    if name == "DBLEngine":
        from .engine import DBLEngine

        return DBLEngine
    raise AttributeError(name)
