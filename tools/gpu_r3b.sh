set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gemm_tc.py -q -m gpu -x -k "dwconv" > gpurun_out/r3b_tests_new.log 2>&1; echo "dwconv tests rc=$?"
tail -6 gpurun_out/r3b_tests_new.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3b_tests_all.log 2>&1; echo "all tests rc=$?"
tail -4 gpurun_out/r3b_tests_all.log
QUICK="--steps 10 --warmup 3 --train-steps 0 --c3-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --parity-utts 8 --ref-utts 4 --ref-utts-max 4"
for v in 1 0 1; do
LFS2_DWCONV_PIPE=$v timeout 600 python bench.py $QUICK > gpurun_out/r3b_bench$v.json 2> gpurun_out/r3b_bench$v.err; echo "bench rc=$?"
python - <<PY
import json
raw=open("gpurun_out/r3b_bench$v.json").read(); d=json.loads(raw[raw.index("{"):])
pk=d["roofline"]["per_kernel"]
print("PIPE=$v ms/step", round(d["ms_per_step"],3), "bf16", round(d["bf16_mode"]["ms_per_step"],3), "padskip", round(d["pad_skip"]["ms_per_step"],3), d["pad_skip"]["valid_frames_bit_identical_to_headline_path"], "parity", d["parity_check"]["c2"]["modes"]["fp32"]["max_abs_mel_err_valid_frames"], d.get("errors"))
for k in ("ffn_fused","lfs2_attention_tc","predictor_pw_ln_gemm","qkv_gemm","out_proj_ln_gemm","lfs2_dwconv1d"): print("  ", k, pk[k]["ms"], pk[k]["frac"])
PY
done
