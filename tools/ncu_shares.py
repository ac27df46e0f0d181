"""Share of one step per kernel function from the ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) -> markdown table.

    python tools/ncu_shares.py profiles/r02_launches.csv [steps]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0  # bench.py --steps 2 --warmup 3
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rd:
        if "gpu__time_duration" not in (r.get("Metric Name") or ""):
            continue
        name = re.sub(r"^void\s+", "", r["Kernel Name"])
        name = re.sub(r"^dbl::", "", name)
        name = re.split(r"[<(]", name)[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        tot[name] += us
        cnt[name] += 1
    allus = sum(tot.values())
    print("| kernel function | launches/step | sum us per step (ncu, serialised) | share |")
    print("|---|---|---|---|")
    for name, us in tot.most_common(26):
        print(f"| `{name}` | {cnt[name] / steps:.1f} | {us / steps:.1f} | {us / allus * 100:.1f}% |")


if __name__ == "__main__":
    main()
