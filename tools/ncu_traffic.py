"""DRAM traffic per launch and kernel family from .ncu-rep files (ncu --set full) -> profiles/rNN_ncu_traffic.json, the file bench.py
reads for `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/prof_*.ncu-rep > profiles/r01_ncu_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

FAMILY = [("dsconv_tc_kernel", "dsconv_tc"),  # before "conv_tc_kernel", which is a substring of it
          ("conv3x3_tc4_kernel", "conv_tc4"), ("conv3x3_tc5_kernel", "conv_tc5"), ("conv3x3s2_tc_kernel", "conv_s2_tc"), ("bottleneck_tc_kernel", "bottleneck_tc"),
          ("stem_tc_kernel", "stem_tc"), ("conv_tc_kernel", "conv_tc_1x1"), ("stem_pair_kernel", "stem_pair"),
          ("bottleneck_mma_kernel", "bottleneck_mma"), ("conv3x3_mma_kernel", "conv_mma"), ("head_decode_kernel", "head_decode"),
          ("dwconv_tile3_kernel", "dwconv"), ("dsconv_mma_kernel", "dsconv"), ("sort_nms_kernel", "sort_nms"),
          ("dysample", "dbl_dysample"), ("gated_add_kernel", "dbl_gated_add")]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    acc = {}
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
        rd = list(csv.reader(io.StringIO(out)))
        if len(rd) < 3:
            continue
        hdr, units = rd[0], rd[1]
        col = {h: i for i, h in enumerate(hdr)}
        scale = {}
        for m in METRICS:
            u = units[col[m]]
            scale[m] = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
        for r in rd[2:]:
            name = r[col["Kernel Name"]]
            fam = next((f for k, f in FAMILY if k in name), None)
            if fam is None:
                continue
            a = acc.setdefault(fam, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += num(r[col["dram__bytes_read.sum"]]) * scale["dram__bytes_read.sum"] + num(r[col["dram__bytes_write.sum"]]) * scale["dram__bytes_write.sum"]
            a[2] += num(r[col["gpu__time_duration.sum"]]) * scale["gpu__time_duration.sum"]
    src = "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches of one bench.py run"
    print(json.dumps({f: {"launches_captured": n, "dram_bytes_per_launch": round(b / n), "time_us_per_launch": round(t / n, 2), "source": src}
                      for f, (n, b, t) in acc.items()}, indent=1))


if __name__ == "__main__":
    main()
