#!/usr/bin/env python
"""Blackwell-native evidence: per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics in liblfs2.so
(cuobjdump -sass).  python tools/sass_counts.py [path/to/lib.so] > profiles/<round>_sass_counts.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lightningfastspeech2_b200", "liblfs2.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2",
             "FFMA2", "FADD2", "FMUL2", "FMNMX3"]   # (the last four: packed fp32 pairs / 3-input maximum of sm_100)
counts = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    for mn in MNEMONICS:
        if re.search(r"\b" + re.escape(mn) + r"\b", line) or (mn.endswith("MMA") and mn in line) or (mn in ("LDTM", "STTM", "SYNCS") and re.search(r"\b" + mn, line)):
            counts[fn][mn] += 1
            break
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("| kernel | " + " | ".join(MNEMONICS) + " |")
print("|---|" + "---:|" * len(MNEMONICS))
tot = collections.Counter()
for (fn, c), name in zip(counts.items(), demangle):
    if not any(c[m] for m in MNEMONICS[:9]):
        continue
    name = re.sub(r"\((int|bool)\)", "", name)
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("lfs2::tc::", "")
    print(f"| `{short}` | " + " | ".join(str(c[m]) for m in MNEMONICS) + " |")
    tot.update(c)
print("| **all kernels with tcgen05/TMA** | " + " | ".join(str(tot[m]) for m in MNEMONICS) + " |")
