"""Import shim for the real reference (player4771/YOLO-DBL) — TEST INFRASTRUCTURE ONLY.

Used only by ``oracle/make_golden.py`` in the authoring container, where
``/root/reference`` is mounted.  It never runs on the GPU box and no product code
imports it.

The reference's ``ultralytics`` fork star-imports three large module zoos
(``nn/tasks.py:101-103``) that need third-party packages which are absent here and
which the YOLO-DBL graph never touches.  We serve inert placeholder packages for
exactly those roots (an allow-list: a catch-all finder would shadow optional
stdlib imports).
"""
import importlib.abc
import importlib.machinery
import os
import sys
import tempfile
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("DBL_REFERENCE_ROOT", "/root/reference")
REF_PKG_DIR = os.path.join(REF_ROOT, "models", "YOLO")

_ABSENT_ROOTS = {
    "matplotlib", "thop", "timm", "pywt", "fairscale", "mmengine", "onnx", "seaborn",
    "torchsummary", "basicsr", "anyup", "albumentations", "pycocotools", "clip",
    "shapely", "mmcv", "mmdet", "natten",
}


class _InertLoader(importlib.abc.Loader):
    def create_module(self, spec):
        mod = MagicMock(name=spec.name)
        mod.__name__ = spec.name
        mod.__path__ = []
        mod.__spec__ = spec
        mod.__loader__ = self
        mod.__version__ = "0.0.0"
        mod.__all__ = []
        return mod

    def exec_module(self, module):
        return None


class _InertFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _ABSENT_ROOTS:
            return importlib.machinery.ModuleSpec(name, _InertLoader(), is_package=True)
        return None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_PKG_DIR, "ultralytics"))


def import_reference():
    """Return the reference's ``ultralytics`` package (CPU)."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    import torch  # noqa: F401  (must precede ultralytics: it pins OMP_NUM_THREADS=1 otherwise)

    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="dbl_ref_cfg_"))
    if not any(isinstance(f, _InertFinder) for f in sys.meta_path):
        sys.meta_path.append(_InertFinder())  # appended last: real packages always win
    if REF_PKG_DIR not in sys.path:
        sys.path.insert(0, REF_PKG_DIR)
    import ultralytics  # noqa: F401

    return ultralytics
