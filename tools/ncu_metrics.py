#!/usr/bin/env python
"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (raw page):
python tools/ncu_metrics.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum [", "dram__bytes_read.sum [", "dram__bytes_write.sum [",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct", "launch__registers_per_thread [",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed [",
        "launch__grid_size", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum [",
        "lts__t_sectors_srcunit_tex_op_write.sum [", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum [",
        "sm__cycles_elapsed.max [", "smsp__inst_executed.sum [", "l1tex__throughput.avg.pct",
        "sm__throughput.avg.pct", "lts__t_bytes.sum ["]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    for v in vals:
        print("KERNEL", v[hdr.index("Kernel Name")][:90])
        for i, h in enumerate(hdr):
            hh = f"{h} [{units[i]}]"
            if any(w in hh for w in WANT) and "Triage" not in hh:
                print(f"  {hh} = {v[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
