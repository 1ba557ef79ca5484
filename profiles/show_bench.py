#!/usr/bin/env python
"""Pretty-print a bench.py JSON line: headline numbers, per-kernel ms, per-layer TFLOP/s."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "clips/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
for k, v in d["kernels"].items():
    if k[0] != "_":
        print(" ", k, v)
    elif k != "_dense_layers_tflops":
        print(" ", k, v)
for k, v in d["kernels"].get("_dense_layers_tflops", {}).items():
    print("   ", k, v)
