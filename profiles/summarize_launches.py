"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
time and share.  Usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/x.txt"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit = r.get("Metric Unit", "ns")
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, v))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    print(f"# {path}: {len(rows)} launches, {tot / 1e3:.3f} ms total (serialised, cold-cache: compare SHARES)")
    print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:60]:60s} {c:8d} {v:12.1f} {v / c:10.2f} {100 * v / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
