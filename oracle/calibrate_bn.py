"""Measure BatchNorm running statistics for the synthetic-weight recipe — TEST/BENCH DATA TOOLING.

    python oracle/calibrate_bn.py            # rewrites yolo_dbl_b200/data/bn_calib_*.npz

A randomly initialised 36-layer residual stack has no stable activation scale (it either collapses onto the BN
biases or overflows fp32, and the turning point moves with width and depth).  Real checkpoints do not have that
problem because BatchNorm's buffers hold the statistics of the training data.  We reproduce exactly that: one forward
pass of the oracle over a seeded batch, every BatchNorm recording the mean/variance of what it sees.  The result is
a small table of per-channel statistics that ``yolo_dbl_b200.synth.synth_state_dict`` overlays on the recipe.
"""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import dbl_oracle as O  # noqa: E402
from yolo_dbl_b200 import synth  # noqa: E402

CASES = [("n", 3, 0), ("n", 80, 0), ("s", 3, 0), ("l", 3, 0), ("n", 80, 3), ("s", 3, 5), ("l", 3, 7)]


def main():
    os.makedirs(synth.CALIB_DIR, exist_ok=True)
    only = sys.argv[1:]
    for scale, nc, seed in CASES:
        if only and f"{scale}{nc}s{seed}" not in only:
            continue
        shapes = json.load(open(os.path.join(ROOT, "tests", "golden", f"shapes_{scale}_nc{nc}.json")))
        sd = synth.synth_state_dict(shapes, seed, calib=None)
        sd["__calibrate__"] = {}
        x = synth.synth_images(4, 640, 640, seed=1234)  # calibrate at the headline resolution
        O.forward(sd, x, scale)
        table = {k: v.numpy() for k, v in sd.pop("__calibrate__").items()}
        # class-logit bias: put ~4 % of the anchors of the calibration batch above conf = 0.25 ("typical" regime)
        _, feats = O.forward(sd, x, scale)
        best = torch.cat([f[:, 64:].flatten(2) for f in feats], 2).amax(1).flatten()
        shift = float(torch.quantile(best[torch.randperm(best.numel(), generator=torch.Generator().manual_seed(0))[:200000]], 0.96)) - math.log(0.25 / 0.75)
        for k in [k for k in sd if ".cv3." in k and k.endswith(".2.bias")]:
            table[k] = (sd[k] - shift).numpy()
        path = synth.calib_path(shapes, seed)
        np.savez_compressed(path, **table)
        y, _ = O.forward(synth.synth_state_dict(shapes, seed), synth.synth_images(1, 640, 640, seed=99), scale)
        frac = float((y[:, 4:].amax(1) > 0.25).float().mean())
        print(f"{os.path.basename(path)}: {len(table)} tensors, {os.path.getsize(path) / 1024:.0f} KiB, "
              f"finite={bool(torch.isfinite(y).all())}, cand>0.25 = {frac * 100:.2f}%")


if __name__ == "__main__":
    main()
