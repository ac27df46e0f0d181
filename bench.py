#!/usr/bin/env python
"""Benchmark of the YOLO-DBL inference hot path on B200 (BASELINE.json metric: images/sec, 640^2, batch 64 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]               # the B200 path (this repo)
    python bench.py --impl reference [--steps K] [--warmup W]         # the reference's algorithm on the host CPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU

A step = forward + Detect decode + NMS over one batch of 64 synthetic 640x640 images per GPU (weak scaling: images are
independent, ranks never exchange data on the path; the only collective is the timing reduction).  One JSON line is
printed by rank 0.  `value` is timed with inputs resident in HBM (CUDA-graph replays); `e2e` goes through the public
API with pinned HOST input and a device->host read of the detections inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images_per_sec_yolo_dbl_n_640_b64_forward_decode_nms"
PEAKS_FALLBACK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
NMS = dict(conf=0.25, iou=0.7, max_det=300)  # predict defaults: cfg/default.yaml:51-53 of the reference


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(PEAKS_FALLBACK, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[1]) for r in rows if len(r) >= 8]
        reasons = set()
        for r in rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][2]) if rows and len(rows[0]) >= 8 else None,
                "power_w_max": max((float(r[3]) for r in rows if len(r) >= 8), default=None), "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (a port of the reference's algorithm; the reference itself is pure Python that
# cannot travel to the GPU box) on all host threads
# ----------------------------------------------------------------------------------------------------------------------
def cpu_port_run(scale, nc, hw, images_per_step, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dbl_oracle as O  # the only place bench.py touches oracle/: as the measured CPU implementation
    import nms_oracle as N
    from yolo_dbl_b200 import param_shapes, resolve, synth_images, synth_state_dict

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = resolve(None, scale, nc)
    sd = synth_state_dict(param_shapes(spec), 0)
    x = synth_images(images_per_step, hw, hw, seed=11)

    def step():
        y, _ = O.forward(sd, x, scale)
        N.non_max_suppression(y.numpy(), NMS["conf"], NMS["iou"], max_det=NMS["max_det"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return images_per_step * steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 32  # bounded sample of the 64-image step: ~0.8 s of CPU work per step on 16 threads
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    ips, ms, cores = cpu_port_run(args.scale, args.nc, args.imgsz, per_step, steps, warmup)
    sample = f"{per_step} images/step of the {args.batch}-image workload, {steps} steps after {warmup} warm-up, oracle port (torch CPU fp32 + C NMS)"
    emit({
        "impl": "reference", "metric": METRIC, "value": round(ips, 3), "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args, 1),
        "cpu_baseline": {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(ips, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def config_dict(args, world):
    return {"workload": f"YOLO-DBL-{args.scale} (yolov13_DBL.yaml, nc={args.nc}) {args.imgsz}x{args.imgsz}, batch {args.batch} per GPU, "
                        f"forward + Detect decode + NMS (conf {NMS['conf']}, iou {NMS['iou']}, max_det {NMS['max_det']})",
            "global_batch": args.batch * world, "per_gpu_batch": args.batch, "imgsz": args.imgsz, "parallelism": f"dp{world} (images sharded, no data-path collective)",
            "weights": "synthetic recipe seed 0 (+BN calibration table)", "l2": "inputs 315 MB/step > 126 MB L2; activations stream through HBM"}


# ----------------------------------------------------------------------------------------------------------------------
def kernel_table(rows, peaks):
    """Aggregate per-launch timings by kernel family and attach the roofline fractions."""
    fam = {}
    for r in rows:
        f = fam.setdefault(r["kern"], dict(kern=r["kern"], launches=0, ms=0.0, flops=0.0, bytes=0.0))
        f["launches"] += 1
        f["ms"] += r["ms"]
        f["flops"] += r["flops"]
        f["bytes"] += r["bytes"]
    total = sum(f["ms"] for f in fam.values()) or 1.0
    out = []
    for f in fam.values():
        s = f["ms"] * 1e-3
        f["share"] = f["ms"] / total
        f["gbs"] = f["bytes"] / s / 1e9 if s else 0.0
        f["tflops"] = f["flops"] / s / 1e12 if s else 0.0
        f["hbm_frac"] = f["gbs"] / peaks["hbm_gbs"]
        f["tensor_frac"] = f["tflops"] / peaks["bf16_tflops_sustained"]
        out.append(f)
    return sorted(out, key=lambda f: -f["ms"])


_STDOUT = sys.stdout


def emit(line: dict):
    """the ONE JSON line on the real stdout (fd 1 itself is pointed at stderr so library banners -- NCCL's version line -- stay out)"""
    print(json.dumps(line), file=_STDOUT, flush=True)


def main():
    global _STDOUT
    _STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", default="n")
    ap.add_argument("--nc", type=int, default=80)
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=2, help="whole 64-image steps in flight per GPU (engine replicas on their own streams)")
    ap.add_argument("--kernel-table", default="", help="write the per-launch timing table (json) here")
    ap.add_argument("--quick", action="store_true", help="device-resident leg only (for runs under ncu): prints a reduced line, never a bench value")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    import torch.distributed as dist

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from yolo_dbl_b200 import DBLBackend, _lib, synth_images
    from yolo_dbl_b200.engine import DBLEngine

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    dev = torch.device("cuda", local)
    model = DBLBackend.synthetic(args.scale, args.nc, seed=0, device=dev)
    eng: DBLEngine = model.engine(args.batch, args.imgsz, args.imgsz)
    small = synth_images(8, args.imgsz, args.imgsz, seed=100 + rank)  # tile 8 distinct images to the batch (CPU synthesis is slow)
    x_host = small.repeat((args.batch + 7) // 8, 1, 1, 1)[: args.batch].contiguous().pin_memory()
    eng.set_input(x_host.to(dev))
    key = eng.post_key(**NMS)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident throughput (CUDA-graph replays of forward + decode + NMS).  `inflight` engine replicas take
    # consecutive steps on their own streams: a step is still one whole pass over one 64-image batch, but step i+1 starts
    # while step i is still running and fills its bubbles (kernel fill/drain, the 64-CTA per-image kernels, small grids).
    P = max(1, args.inflight)
    engs = [eng] + [model.engine(args.batch, args.imgsz, args.imgsz, replica=r) for r in range(1, P)]
    for e in engs[1:]:
        e.set_input(x_host.to(dev))
    side = [torch.cuda.Stream(device=dev) for _ in range(P)]
    step = lambda: eng.detect(None, **NMS)  # noqa: E731

    def run_steps(k):
        main = torch.cuda.current_stream(dev)
        for s in side:
            s.wait_stream(main)
        for i in range(k):
            with torch.cuda.stream(side[i % P]):
                engs[i % P].detect(None, **NMS)
        for s in side:
            main.wait_stream(s)

    clk = ClockSampler(local) if rank == 0 and not args.quick else None  # samples SM clocks / throttle reasons every 100 ms until the e2e legs are done
    run_steps(max(warmup, P))
    ms_total = timed(lambda: run_steps(steps), 1)
    value = args.batch * world * steps / (ms_total * 1e-3)
    kps = eng.kernels_per_step.get((key, False, False), 0)
    launches = kps * steps

    if args.quick:
        if rank == 0:
            emit({"metric": METRIC, "quick": True, "value": round(value, 1), "ms_per_step": round(ms_total / steps, 4), "gpu_launches": launches})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public API: pinned host batches -> H2D -> predict -> D2H of the detections, every step.
    # Image sources reach the reference as uint8 and are scaled on the device (engine/predictor.py:116-134); the same
    # here: uint8 NCHW crosses PCIe, the /255 is fused into the stem kernel.  detect_stream() overlaps the copy of batch
    # i+1 with the compute of batch i (the reference's predict(stream=True) generator).
    out_rows = torch.empty(args.batch, NMS["max_det"], 6, pin_memory=True)
    out_cnt = torch.empty(args.batch, dtype=torch.int32, pin_memory=True)
    x_u8 = [(x_host * 255.0).round().to(torch.uint8).pin_memory() for _ in range(2)]

    from yolo_dbl_b200.parallel import gather_detections

    d2h = torch.cuda.Stream(device=dev)

    def e2e_run(k, src):
        n = 0
        for rows, counts, ev in model.detect_stream((src[i & 1] for i in range(k)), with_event=True, in_flight=P, **NMS):
            d2h.wait_event(ev)  # this batch's result is complete; the next batch is already queued on the compute stream
            with torch.cuda.stream(d2h):
                if world > 1:  # the path's one collective: final gather of the fixed-capacity detections (NCCL all-gather)
                    g_rows, g_cnt = gather_detections(rows, counts, args.batch * world)
                    rows, counts = g_rows[rank * args.batch:(rank + 1) * args.batch], g_cnt[rank * args.batch:(rank + 1) * args.batch]
                out_rows.copy_(rows, non_blocking=True)
                out_cnt.copy_(counts, non_blocking=True)
            d2h.synchronize()  # the caller reads the result on the host every step
            n += 1
        return n

    e2e_steps = max(3, steps)
    e2e_run(3, x_u8)
    e2e_ms = timed(lambda: e2e_run(e2e_steps, x_u8), 1)
    e2e_value = args.batch * world * e2e_steps / (e2e_ms * 1e-3)
    detections = int(out_cnt.sum())
    x_f32 = [x_host, x_host.clone().pin_memory()]
    e2e_run(2, x_f32)
    e2e_f32_ms = timed(lambda: e2e_run(e2e_steps, x_f32), 1)
    e2e_f32_value = args.batch * world * e2e_steps / (e2e_f32_ms * 1e-3)

    if clk is not None:  # keep the GPU under the same load until nvidia-smi has a handful of samples (the timed legs last < 0.5 s)
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            step()
        torch.cuda.synchronize(dev)
    clocks = clk.stop() if clk else None

    # ---- per-kernel roofline table (rank 0, eager, each launch between its own CUDA events)
    peaks = load_peaks()
    roof, table = None, []
    if rank == 0:
        rows = eng.profile_ops(reps=5, post_key=key)
        table = kernel_table(rows, peaks)
        top = table[0]  # the kernel (family) with the largest share of the step; per-launch averages below
        t_tensor, t_hbm = top["flops"] / (peaks["bf16_tflops_sustained"] * 1e12), top["bytes"] / (peaks["hbm_gbs"] * 1e9)
        bound = "tensor" if t_tensor > t_hbm else "hbm"
        ach, peak, unit = (top["tflops"], peaks["bf16_tflops_sustained"], "TFLOP/s") if bound == "tensor" else (top["gbs"], peaks["hbm_gbs"], "GB/s")
        traffic = None  # DRAM bytes per launch from the committed ncu capture of the same command (profiles/r01_ncu_traffic.json)
        tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(top["kern"], {}).get("dram_bytes_per_launch")
        n = max(1, top["launches"])
        roof = {"kernel": top["kern"], "bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": unit, "frac": round(ach / peak, 4),
                "traffic": traffic, "launches_per_step": top["launches"], "share_of_step": round(top["share"], 3),
                "avg_launch_us": round(top["ms"] * 1e3 / n, 2), "algorithmic_bytes_per_launch": round(top["bytes"] / n),
                "algorithmic_flops_per_launch": round(top["flops"] / n), "peak_source": peaks["source"],
                "tensor_frac": round(top["tflops"] / peaks["bf16_tflops_sustained"], 4), "hbm_frac": round(top["gbs"] / peaks["hbm_gbs"], 4),
                "note": "achieved = sum of algorithmic bytes (or flops) of this kernel's launches in one step / sum of their CUDA-event times; "
                        "per-launch rows: --kernel-table"}
        if args.kernel_table:
            os.makedirs(os.path.dirname(os.path.abspath(args.kernel_table)), exist_ok=True)
            json.dump({"families": table, "launches": rows, "peaks": peaks}, open(args.kernel_table, "w"), indent=1)

    # ---- CPU baseline (rank 0, N == 1 only): the oracle port on all host threads, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, ms, cores = cpu_port_run(args.scale, args.nc, args.imgsz, 64, 6, 1)
        cpu = {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"384 images (6 steps x 64 = the full {args.batch}-image step) after 1 warm-up step, ~10 s of CPU work; oracle port: "
                         "torch CPU fp32 forward + C NMS on all host threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": round(ms_total / steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": dict(config_dict(args, world), steps_in_flight=P), "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": x_u8[0].numel(),
                    "d2h_bytes_per_step": out_rows.numel() * 4 + out_cnt.numel() * 4, "ms_per_step": round(e2e_ms / e2e_steps, 4),
                    "input": "uint8 NCHW in pinned host memory (image source; /255 fused into the stem kernel), copy of batch i+1 "
                             "overlapped with compute of batch i, detections read back to pinned host memory every step",
                    "fp32_input": {"value": round(e2e_f32_value, 1), "h2d_bytes_per_step": x_host.numel() * 4,
                                   "ms_per_step": round(e2e_f32_ms / e2e_steps, 4)}},
            "gpu_launches": launches, "kernels_per_step": kps, "detections_per_step": detections,
            "roofline": roof, "cpu_baseline": cpu,
            "kernels": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in f.items() if k in ("kern", "launches", "ms", "share", "gbs", "tflops", "hbm_frac", "tensor_frac")}
                        for f in table[:8]],
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
