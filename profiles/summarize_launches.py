#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total
device time, launch count and share.  Usage: summarize_launches.py launches.csv > summary.md"""
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
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"<.*", "", name) if "at::native" in name else name
        rows.append((name, val * scale))
    agg = defaultdict(lambda: [0.0, 0])
    for n, ms in rows:
        agg[n][0] += ms
        agg[n][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {total:.2f} ms total device time (cold-cache, serialised)\n")
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for n, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"| `{n[:110]}` | {c} | {ms:.3f} | {100 * ms / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
