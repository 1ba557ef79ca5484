#!/usr/bin/env python
"""Key metrics of an `ncu --set full` capture as a markdown table.
Usage: ncu_summary.py file.ncu-rep > summary.md   (needs `ncu` on PATH; no GPU)"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## `{name[:100]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
