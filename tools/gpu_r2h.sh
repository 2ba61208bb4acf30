set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_forward.py -q -m gpu > gpurun_out/r2h_tests_attn.log 2>&1; echo "attn tests rc=$?"
tail -5 gpurun_out/r2h_tests_attn.log
QUICK="--steps 10 --warmup 3 --train-steps 0 --c3-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --buckets --parity-utts 8 --ref-utts 4 --ref-utts-max 4"
LFS2_ATTN_PP=0 timeout 600 python bench.py $QUICK > gpurun_out/r2h_bench_pp0.json 2> gpurun_out/r2h_bench_pp0.err; echo "bench pp0 rc=$?"
LFS2_ATTN_PP=1 timeout 600 python bench.py $QUICK > gpurun_out/r2h_bench_pp1.json 2> gpurun_out/r2h_bench_pp1.err; echo "bench pp1 rc=$?"
python - <<'PY'
import json
for n in ("pp0","pp1"):
    raw=open(f"gpurun_out/r2h_bench_{n}.json").read(); d=json.loads(raw[raw.index("{"):])
    pk=d["roofline"]["per_kernel"]
    print(n, "ms/step", round(d["ms_per_step"],3), "attn ms(3 steps)", pk["lfs2_attention_tc"]["ms"], "tflops", pk["lfs2_attention_tc"]["tflops"], "bf16 ms", round(d["bf16_mode"]["ms_per_step"],3), "padskip ms", round(d["pad_skip"]["ms_per_step"],3), "parity", d["parity_check"]["c2"]["modes"]["fp32"]["max_abs_mel_err_valid_frames"], d.get("errors"))
PY
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2h_tests_all.log 2>&1; echo "all tests rc=$?"
tail -4 gpurun_out/r2h_tests_all.log
