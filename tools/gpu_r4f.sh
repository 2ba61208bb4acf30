set -x
mkdir -p gpurun_out
M="sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_bytes.sum"
timeout 900 ncu --set full --metrics $M --clock-control none --cache-control none -k regex:"gemm_tc2_kernel" -s 36 -c 14 -f -o /tmp/r4f_g2 python tools/profile_train.py 2 fp32 C4 > gpurun_out/r4f_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_extract.py /tmp/r4f_g2.ncu-rep > gpurun_out/r4f_gemm_tc2_ncu.txt 2>&1
grep -c KERNEL gpurun_out/r4f_gemm_tc2_ncu.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 900 --csv --log-file gpurun_out/r4f_train_launches.csv python tools/profile_train.py 2 fp32 C4 > gpurun_out/r4f_ncu2.log 2>&1; echo "ncu list rc=$?"
