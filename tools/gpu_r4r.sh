set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"dwconv1d_k_kernel|dwconv1d_tma_kernel" -s 6 -c 8 -f -o /tmp/r4r_dw python tools/dwconv_ab.py > gpurun_out/r4r_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r4r_dw.ncu-rep --page raw --csv > /tmp/r4r_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/r4r_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2:]
want=[h for h in hdr if any(k in h for k in ('Kernel Name','gpu__time_duration.sum','sm__throughput.avg.pct','smsp__issue_active.avg.pct','sm__inst_executed_pipe_fma','sm__pipe_fma_cycles_active.avg.pct','sm__pipe_fmaheavy','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__average_warp','warps_issue_stalled','smsp__inst_executed.sum','achieved_occupancy','sm__warps_active','dram__throughput','lts__throughput','l1tex__throughput','launch__registers','launch__occupancy_limit'))]
out=open('gpurun_out/r4r_dwconv_ncu.txt','w')
for v in vals:
    for h in want:
        i=hdr.index(h)
        out.write(f"{h} [{units[i]}] = {v[i]}\n")
    out.write("\n")
out.close()
PY
grep -c . gpurun_out/r4r_dwconv_ncu.txt
