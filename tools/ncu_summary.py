"""Summarise .ncu-rep files (ncu --set full) into a small markdown table: duration, DRAM bytes, pipe utilisation, issue rate.

    python tools/ncu_summary.py gpurun_out/prof_tc4_r1.ncu-rep [more.ncu-rep ...] > profiles/r01_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr = rd[0]
    for r in rd[2:]:
        yield dict(zip(hdr, r))


def f(d, k, scale=1.0, fmt="{:.1f}"):
    v = d.get(k, "")
    try:
        return fmt.format(float(v.replace(",", "")) * scale)
    except ValueError:
        return "-"


def main():
    print("| kernel | grid x block | regs | time us | DRAM rd MB | DRAM wr MB | SM thr % | issue % | warps % | tensor(hmma) % | L2 hit % | warp-inst M |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for path in sys.argv[1:]:
        print(f"| **{path.split('/')[-1]}** | | | | | | | | | | | |")
        for d in rows_of(path):
            name = d["Kernel Name"].replace("void ", "").replace("dbl::", "")
            name = name[:name.index("(")] if "(" in name else name
            hm = f(d, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")
            if hm == "-":
                hm = f(d, "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active")
            print(f"| `{name[:60]}` | {f(d, 'launch__grid_size', fmt='{:.0f}')} x {f(d, 'launch__block_size', fmt='{:.0f}')} | "
                  f"{f(d, 'launch__registers_per_thread', fmt='{:.0f}')} | {f(d, 'gpu__time_duration.sum')} | {f(d, 'dram__bytes_read.sum')} | "
                  f"{f(d, 'dram__bytes_write.sum')} | {f(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed')} | "
                  f"{f(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active')} | {f(d, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | "
                  f"{hm} | {f(d, 'lts__t_sector_hit_rate.pct')} | {f(d, 'smsp__inst_executed.sum', 1e-6, '{:.2f}')} |")


if __name__ == "__main__":
    main()
