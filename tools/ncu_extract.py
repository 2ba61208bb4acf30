#!/usr/bin/env python
"""Selected metrics of every kernel in an .ncu-rep (raw page) as a table:
python tools/ncu_extract.py file.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2:]
        print(f"## {path}")
        for v in vals:
            print("KERNEL", v[hdr.index("Kernel Name")][:100], "grid", v[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
            for w in WANT:
                for i, h in enumerate(hdr):
                    if h == w:
                        print(f"  {h} [{units[i]}] = {v[i]}")


if __name__ == "__main__":
    main(sys.argv[1:])
