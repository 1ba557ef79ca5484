#!/usr/bin/env python
"""profiles/traffic.json (read by bench.py for `roofline.traffic`) from an ncu launch list: mean DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum) of the kernels behind each C-ABI entry point.
Usage: python profiles/make_traffic.py profiles/r2_launches_final.csv.gz > profiles/traffic.json"""
import csv
import gzip
import json
import sys
from collections import defaultdict

ENTRY = {"conv_tc_fwd_halo2_kernel": "tag_conv_tc_fwd_halo", "conv_tc_fwd_halo_kernel": "tag_conv_tc_fwd_halo",
         "logmel_warp_kernel": "tag_logmel_fwd_v2", "conv_tc_wgrad2_kernel": "tag_conv_tc_wgrad",
         "conv_tc_wgrad_kernel": "tag_conv_tc_wgrad", "conv_tc_wgrad64_kernel": "tag_conv_tc_wgrad64"}


def main(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt", newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    per_launch = defaultdict(lambda: defaultdict(float))      # (entry, launch id) -> metric -> value
    for r in csv.DictReader(lines):
        name = r["Kernel Name"]
        entry = next((e for k, e in ENTRY.items() if k in name), None)
        if entry is None or r["Metric Name"] not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "byte")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per_launch[(entry, r["ID"])][r["Metric Name"]] += val * scale
    agg = defaultdict(list)
    for (entry, _), m in per_launch.items():
        agg[entry].append(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"])
    out = {e: {"traffic_bytes_per_launch": sum(v) / len(v), "launches": len(v),
               "source": f"{path}: mean of dram__bytes_read.sum + dram__bytes_write.sum over the launches of the kernels "
                         "behind this entry point (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                         "dram__bytes_write.sum --clock-control none, bench.py --steps 2 --warmup 1 --settle 0, scripts/final_1gpu.sh)"}
           for e, v in agg.items()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
