"""Builds a file with the SAME pickle structure as a reference ``.pt`` checkpoint -- a module tree whose classes live under
``ultralytics.nn.*`` -- without the reference package: the classes are defined in throw-away modules that are removed from
``sys.modules`` again before the file is read, exactly the situation on a box that only has yolo_dbl_b200."""
import sys
import types
from collections import OrderedDict

import torch
import torch.nn as nn

_FAKE = ("ultralytics", "ultralytics.nn", "ultralytics.nn.tasks", "ultralytics.nn.modules", "ultralytics.nn.modules.block")


def write_fake_reference_checkpoint(path, state_dict, yaml_dict, names, half=True, use_ema=True):
    mods = {n: types.ModuleType(n) for n in _FAKE}
    sys.modules.update(mods)
    try:
        Node = type("C3k2", (nn.Module,), {"__module__": "ultralytics.nn.modules.block"})
        Model = type("DetectionModel", (nn.Module,), {"__module__": "ultralytics.nn.tasks"})
        mods["ultralytics.nn.modules.block"].C3k2 = Node
        mods["ultralytics.nn.tasks"].DetectionModel = Model
        root = Model()
        for key, t in state_dict.items():
            parts = key.split(".")
            node = root
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, Node())
                node = node._modules[p]
            t = t.clone()
            if parts[-1] in ("running_mean", "running_var", "num_batches_tracked", "init_pos"):
                node.register_buffer(parts[-1], t)
            else:
                node.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
        root.yaml, root.names, root.nc, root.task = dict(yaml_dict), dict(names), yaml_dict.get("nc"), "detect"
        assert list(root.state_dict().keys()) == list(state_dict.keys())
        if half:
            root = root.half()
        ck = {"epoch": 7, "best_fitness": None, "model": None if use_ema else root, "ema": root if use_ema else None, "updates": 1,
              "optimizer": None, "train_args": {"imgsz": 640, "model": "yolov13n_DBL.yaml"}, "date": "2026-01-01", "version": "8.3.63"}
        torch.save(ck, path)
    finally:
        for n in _FAKE:
            sys.modules.pop(n, None)
    return OrderedDict((k, (v.half().float() if (half and v.is_floating_point()) else v)) for k, v in state_dict.items())
