set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2f_tests_all.log 2>&1; echo "all tests rc=$?"
tail -6 gpurun_out/r2f_tests_all.log
timeout 300 python tools/profile_c3.py bf16 32 > gpurun_out/r2f_c3_profile_bf16.txt 2>&1; head -16 gpurun_out/r2f_c3_profile_bf16.txt
timeout 300 python tools/profile_vocoder.py 8 --table > gpurun_out/r2f_vocoder_profile.txt 2>&1; head -14 gpurun_out/r2f_vocoder_profile.txt
M="smsp__inst_executed_pipe_uniform.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum"
timeout 600 ncu --set full --metrics $M --clock-control none --cache-control none --import-source on -k regex:"attention_tc_kernel|ffn_fused_tc_kernel" -s 24 -c 2 -f -o gpurun_out/r2f_c2_fp32 python tools/profile_step.py 3 fp32 > gpurun_out/r2f_ncu1.log 2>&1
timeout 600 ncu --set full --metrics $M --clock-control none --cache-control none --import-source on -k regex:"attention_tc_wide_kernel" -s 14 -c 1 -f -o gpurun_out/r2f_c3_wide python tools/profile_c3.py bf16 32 > gpurun_out/r2f_ncu2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2f_synth_launches.csv python tools/profile_step.py 4 fp32 > gpurun_out/r2f_ncu3.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2f_bench_n1.err
