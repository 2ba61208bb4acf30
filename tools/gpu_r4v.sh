set -x
mkdir -p gpurun_out
M="smsp__inst_executed_pipe_uniform.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum"
timeout 900 ncu --set full --metrics $M --clock-control none --cache-control none -k regex:"ffn_fused_tc_kernel|attention_tc_pp_kernel|gemm_tc_kernel|dwconv1d" -s 120 -c 60 -f -o /tmp/r4v_c2_step python tools/profile_step.py 3 fp32 > gpurun_out/r4v_ncu1.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_extract.py /tmp/r4v_c2_step.ncu-rep > gpurun_out/r4v_c2_step_ncu.txt 2>&1
grep -c KERNEL gpurun_out/r4v_c2_step_ncu.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r4v_synth_launches.csv python tools/profile_step.py 5 fp32 > gpurun_out/r4v_ncu2.log 2>&1; echo "ncu list rc=$?"
