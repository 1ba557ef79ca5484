#!/usr/bin/env python
"""Blackwell-native instruction evidence: per kernel of the built libtag_b200.so, the count of the SASS mnemonics that prove
tcgen05 / TMEM / TMA / CTA-pair / st.async / 256-bit use (B200_PROFILING.md, "What proves a Blackwell-native kernel").
Usage: python profiles/sass_evidence.py > profiles/r2_sass_evidence.md      (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "texttoaudiogrounding_b200", "lib", "libtag_b200.so")
PATTERNS = [("UTCHMMA.2CTA", r"UTCHMMA\.2CTA"), ("UTCHMMA", r"UTCHMMA(?!\.2CTA)"), ("UTMALDG.*.2CTA", r"UTMALDG\.[0-9A-Z]+\.2CTA"),
            ("UTMALDG", r"UTMALDG\.[0-9A-Z]+(?!\.2CTA)\b"), ("UTCBAR.2CTA.MULTICAST", r"UTCBAR\.2CTA\.MULTICAST"),
            ("UTCBAR", r"UTCBAR(?!\.2CTA)"), ("LDTM", r"\bLDTM"), ("UCGABAR", r"UCGABAR_(ARV|WAIT)"), ("STAS (st.async)", r"\bSTAS"),
            ("HMMA (mma.sync)", r"\bHMMA"), ("MOVM (movmatrix)", r"\bMOVM"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
            ("STG/LDG .256", r"\b(STG|LDG)\.E\.[A-Z0-9.]*256"), ("MUFU.TANH", r"MUFU\.TANH")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"^void ", "", name)
            cur = re.sub(r"\(.*", "", name)
            counts.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        for label, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][label] += 1
    print("# SASS evidence (`cuobjdump -sass texttoaudiogrounding_b200/lib/libtag_b200.so`, sm_100a)\n")
    print("Counts of instructions per kernel; kernels without any of them (plain elementwise / reduction kernels) are omitted.\n")
    labels = [l for l, _ in PATTERNS]
    print("| kernel | " + " | ".join(labels) + " |")
    print("|---|" + "---:|" * len(labels))
    for k, c in counts.items():
        if not c:
            continue
        print(f"| `{k[:110]}` | " + " | ".join(str(c.get(l, "")) for l in labels) + " |")


if __name__ == "__main__":
    main()
