set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3j_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r3j_bench_n1.json 2> gpurun_out/r3j_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3j_bench_reference.json 2> gpurun_out/r3j_bench_reference.err; echo "ref rc=$?"
M="smsp__inst_executed_pipe_uniform.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum"
timeout 900 ncu --set full --metrics $M --clock-control none --cache-control none -k regex:"ffn_fused_tc_kernel|attention_tc_pp_kernel|gemm_tc_kernel" -s 92 -c 46 -f -o /tmp/r3j_c2_step python tools/profile_step.py 3 fp32 > gpurun_out/r3j_ncu1.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_extract.py /tmp/r3j_c2_step.ncu-rep > gpurun_out/r3j_c2_step_ncu.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r3j_synth_launches.csv python tools/profile_step.py 5 fp32 > gpurun_out/r3j_ncu2.log 2>&1; echo "ncu list rc=$?"
du -sh gpurun_out
