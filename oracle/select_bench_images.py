"""Choose the synthetic images bench.py measures on — TEST/BENCH DATA TOOLING (authoring container, CPU oracle).

    python oracle/select_bench_images.py        # prints the seed list to paste into yolo_dbl_b200/synth.py:BENCH_IMAGE_SEEDS

The synthetic weights are a random deep stack: on some images the fp32 network itself is ill-conditioned (activations grow to
|x| ~ 60 by layer 34, 10-40 % of all anchors pass conf = 0.25, and the reference's OWN forward with bf16 activation storage then
moves boxes by tens of pixels) -- nothing can be compared on those.  SURVEY.md §8d defines the headline workload as the TYPICAL
regime (about 1-3 % of anchors above conf = 0.25).  Rule applied here, with the CPU oracle only: walk seeds 100, 101, ... and keep
image synth_images(1, 640, 640, seed) when (a) 0.5 % <= candidates <= 3 % of the 8400 anchors and (b) the oracle's bf16-emulated
forward stays within 0.5 px (p99 of |d box|) and 0.02 (max |d score|) of its fp32 forward.  The first 8 such seeds are the batch.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import dbl_oracle as O  # noqa: E402
from yolo_dbl_b200 import param_shapes, resolve, synth_images, synth_state_dict  # noqa: E402


def main(n=8, scale="n", nc=80, wseed=0):
    sd = synth_state_dict(param_shapes(resolve(None, scale, nc)), wseed)
    keep, seed = [], 100
    while len(keep) < n:
        x = synth_images(1, 640, 640, seed)
        y, _ = O.forward(sd, x, scale)
        cand = int((y[:, 4:].amax(1) > 0.25).sum())
        if 42 <= cand <= 252:
            y2, _ = O.forward(sd, x, scale, round_bf16=True)
            d = (y - y2).abs()
            p99, smax = float(d[:, :4].flatten().quantile(0.99)), float(d[:, 4:].max())
            ok = p99 <= 0.5 and smax <= 0.02
            print(f"seed {seed}: {cand} candidates, bf16 drift box p99 {p99:.3f} px, score max {smax:.4f} -> {'keep' if ok else 'skip'}")
            if ok:
                keep.append(seed)
        else:
            print(f"seed {seed}: {cand} candidates -> outside the typical regime")
        seed += 1
    print("BENCH_IMAGE_SEEDS =", tuple(keep))


if __name__ == "__main__":
    main()
