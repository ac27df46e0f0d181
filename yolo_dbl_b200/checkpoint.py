"""Ingestion of the reference's ``.pt`` checkpoints without the reference package.

The reference saves ``{"model": <pickled DetectionModel>, "ema": ..., "train_args": ...}`` (``engine/trainer.py`` save_model) and loads it
by unpickling the module tree with its own classes importable (``nn/tasks.py:804-944``: ``torch_safe_load`` ->
``attempt_load_one_weight``: ``(ckpt.get("ema") or ckpt["model"]).float()``).  On a box that only has this package those classes do
not exist, so the unpickler here resolves every class outside ``torch`` / ``collections`` / builtins to an inert placeholder: the
``nn.Module`` state of each node (``_parameters``, ``_buffers``, ``_modules``) is plain data and is all that is needed to rebuild the
``state_dict`` -- same keys, same order -- plus the model's ``yaml`` / ``names`` / ``nc`` attributes.  No code from the checkpoint runs.
"""
from __future__ import annotations

import io
import pickle
import types
from collections import OrderedDict
from typing import Dict, Tuple

import torch

_ALLOWED_ROOTS = ("torch", "collections", "builtins", "__builtin__", "numpy", "_codecs", "copyreg", "copy_reg", "pathlib", "argparse")


class _Placeholder:
    """stands in for every class the checkpoint names that is not importable here (ultralytics.nn.modules.*, ...)"""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:  # (dict, slots) form
            for part in state if isinstance(state, tuple) else (state,):
                if isinstance(part, dict):
                    self.__dict__.update(part)

    def __call__(self, *a, **k):
        return self


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        root = module.split(".")[0]
        if root in _ALLOWED_ROOTS:
            try:
                return super().find_class(module, name)  # incl. the protocol-2 spellings (__builtin__, copy_reg) torch.save emits
            except (ImportError, AttributeError):
                pass
        # a distinct placeholder class per (module, name) so isinstance / repr stay meaningful
        return type(name, (_Placeholder,), {"__module__": module})


_pickle_module = types.ModuleType("dbl_checkpoint_pickle")
_pickle_module.Unpickler = _Unpickler
_pickle_module.load = lambda f, **kw: _Unpickler(f, **{k: v for k, v in kw.items() if k == "encoding"}).load()
_pickle_module.__name__ = "pickle"


def _collect(node, prefix: str, out: "OrderedDict[str, torch.Tensor]"):
    """nn.Module.state_dict() order: own parameters, own persistent buffers, then children in registration order"""
    d = getattr(node, "__dict__", {})
    for k, v in (d.get("_parameters") or {}).items():
        if v is not None:
            out[prefix + k] = v.detach() if isinstance(v, torch.Tensor) else v
    non_persistent = d.get("_non_persistent_buffers_set") or set()
    for k, v in (d.get("_buffers") or {}).items():
        if v is not None and k not in non_persistent:
            out[prefix + k] = v
    for k, child in (d.get("_modules") or {}).items():
        if child is not None:
            _collect(child, prefix + k + ".", out)


def load_pt(path_or_file) -> Tuple[Dict[str, torch.Tensor], dict]:
    """Returns (state_dict in fp32, info) with info = {yaml, names, nc, scale, task, args}.

    Follows ``attempt_load_one_weight`` (reference ``nn/tasks.py:897-924``): the EMA weights when present, else ``model``; ``.float()``.
    A bare ``state_dict`` file (``torch.save(model.state_dict())``) is accepted too."""
    if isinstance(path_or_file, (str, bytes)) or hasattr(path_or_file, "__fspath__"):
        with open(path_or_file, "rb") as f:
            blob = f.read()
    else:
        blob = path_or_file.read()
    ckpt = torch.load(io.BytesIO(blob), map_location="cpu", weights_only=False, pickle_module=_pickle_module)
    info = {"yaml": None, "names": None, "nc": None, "scale": None, "task": "detect", "args": {}}
    if isinstance(ckpt, dict) and ("model" in ckpt or "ema" in ckpt):
        model = ckpt.get("ema") or ckpt["model"]
        info["args"] = ckpt.get("train_args") or {}
    elif isinstance(ckpt, dict) and all(isinstance(v, torch.Tensor) for v in ckpt.values()):
        return OrderedDict((k, v.float() if v.is_floating_point() else v) for k, v in ckpt.items()), info
    else:
        model = ckpt
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _collect(model, "", sd)
    if not sd:
        raise ValueError("no parameters found in the checkpoint: not an ultralytics-style model file")
    sd = OrderedDict((k, v.float() if v.is_floating_point() else v) for k, v in sd.items())
    y = getattr(model, "yaml", None)
    if isinstance(y, dict):
        info["yaml"], info["scale"], info["nc"] = y, y.get("scale"), y.get("nc")
    names = getattr(model, "names", None)
    if names is not None:
        info["names"] = dict(names) if isinstance(names, dict) else {i: n for i, n in enumerate(names)}
    if info["nc"] is None:
        ks = [v.shape[0] for k, v in sd.items() if ".cv3." in k and k.endswith(".2.weight")]
        info["nc"] = ks[0] if ks else None
    return sd, info
