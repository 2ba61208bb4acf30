set -x
M="smsp__inst_executed_pipe_uniform.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum"
timeout 600 ncu --set full --metrics $M --clock-control none --cache-control none --import-source on -k regex:attention_tc_kernel -s 12 -c 2 -f -o gpurun_out/r2a_attention_fp32 python tools/profile_step.py 3 fp32 > gpurun_out/r2a_ncu1.log 2>&1
timeout 600 ncu --set full --metrics $M --clock-control none --cache-control none --import-source on -k regex:ffn_fused_tc_kernel -s 12 -c 2 -f -o gpurun_out/r2a_ffn_fp32 python tools/profile_step.py 3 fp32 > gpurun_out/r2a_ncu2.log 2>&1
timeout 600 ncu --set full --metrics $M --clock-control none --cache-control none --import-source on -k regex:"attention_tc_kernel|ffn_fused_tc_kernel" -s 24 -c 2 -f -o gpurun_out/r2a_bf16 python tools/profile_step.py 3 bf16 > gpurun_out/r2a_ncu3.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
tail -3 gpurun_out/r2a_ncu1.log gpurun_out/r2a_ncu2.log gpurun_out/r2a_ncu3.log
