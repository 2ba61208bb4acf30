#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launch count, total / mean duration and share of the step (shares, not absolutes, are
what carries over to an un-profiled run: ncu serialises launches and runs cold-cache)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, us in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total us | mean us | share |")
    print(f"|---|---:|---:|---:|---:|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {total:.1f} | | 100% |")


if __name__ == "__main__":
    main(sys.argv[1])
