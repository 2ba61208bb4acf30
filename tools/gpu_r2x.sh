set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_round2.py -q -m gpu -x -s -k "layernorm or multicast or fused_predictor or residual" > gpurun_out/r2x_tests_new.log 2>&1; echo "new tests rc=$?"
grep -n "two-pass\|passed\|failed\|^E  " gpurun_out/r2x_tests_new.log | head -30
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2x_tests_all.log 2>&1; echo "all tests rc=$?"
tail -15 gpurun_out/r2x_tests_all.log
QUICK="--steps 10 --warmup 3 --train-steps 0 --c3-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --buckets --parity-utts 8 --ref-utts 4 --ref-utts-max 4"
timeout 600 python bench.py $QUICK > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
raw=open("gpurun_out/r2x_bench.json").read(); d=json.loads(raw[raw.index("{"):])
pk=d["roofline"]["per_kernel"]
print("ms/step", round(d["ms_per_step"],3), "bf16", round(d["bf16_mode"]["ms_per_step"],3), "padskip", round(d["pad_skip"]["ms_per_step"],3), d["pad_skip"]["valid_frames_bit_identical_to_headline_path"], "parity", d["parity_check"]["c2"]["modes"]["fp32"], d.get("errors"))
for k,v in pk.items(): print(k, v)
PY
