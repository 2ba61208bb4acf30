set -x
mkdir -p gpurun_out
LFS2_ATTN_PP=2 timeout 600 python -m pytest tests/test_gpu_attention_tc.py -q -m gpu > gpurun_out/r2i_tests_attn_pp2.log 2>&1; echo "attn tests pp2 rc=$?"
tail -3 gpurun_out/r2i_tests_attn_pp2.log
QUICK="--steps 10 --warmup 3 --train-steps 0 --c3-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --buckets --parity-utts 8 --ref-utts 4 --ref-utts-max 4"
for v in 0 1 2 1 2; do
LFS2_ATTN_PP=$v timeout 600 python bench.py $QUICK > gpurun_out/r2i_bench_pp$v.json 2> gpurun_out/r2i_bench_pp$v.err; echo "bench pp$v rc=$?"
python - <<PY
import json
n="pp$v"
raw=open(f"gpurun_out/r2i_bench_{n}.json").read(); d=json.loads(raw[raw.index("{"):])
pk=d["roofline"]["per_kernel"]
print(n, "ms/step", round(d["ms_per_step"],3), "attn ms(3 steps)", pk["lfs2_attention_tc"]["ms"], "tflops", pk["lfs2_attention_tc"]["tflops"], "bf16 ms", round(d["bf16_mode"]["ms_per_step"],3), "padskip ms", round(d["pad_skip"]["ms_per_step"],3), "parity", d["parity_check"]["c2"]["modes"]["fp32"]["max_abs_mel_err_valid_frames"], d.get("errors"))
PY
done
M="smsp__inst_executed_pipe_uniform.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum"
LFS2_ATTN_PP=1 timeout 600 ncu --set full --metrics $M --clock-control none --cache-control none --import-source on -k regex:"attention_tc_pp_kernel" -s 12 -c 1 -f -o gpurun_out/r2i_attn_pp1 python tools/profile_step.py 3 fp32 > gpurun_out/r2i_ncu1.log 2>&1
LFS2_ATTN_PP=2 timeout 600 ncu --set full --metrics $M --clock-control none --cache-control none --import-source on -k regex:"attention_tc_wide_kernel" -s 12 -c 1 -f -o gpurun_out/r2i_attn_pp2 python tools/profile_step.py 3 fp32 > gpurun_out/r2i_ncu2.log 2>&1
