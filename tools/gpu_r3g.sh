set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention_tc.py -q -m gpu -x > gpurun_out/r3g_tests_new.log 2>&1; echo "attn tests rc=$?"
tail -4 gpurun_out/r3g_tests_new.log
python tools/attn_timeline.py run 2>&1 | tee gpurun_out/r3g_attn_timeline.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3g_tests_all.log 2>&1; echo "all tests rc=$?"
tail -3 gpurun_out/r3g_tests_all.log
QUICK="--steps 10 --warmup 3 --train-steps 0 --c3-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --parity-utts 8 --ref-utts 4 --ref-utts-max 4"
for i in 1 2; do
timeout 600 python bench.py $QUICK > gpurun_out/r3g_bench$i.json 2> gpurun_out/r3g_bench$i.err; echo "bench rc=$?"
python - <<PY
import json
raw=open("gpurun_out/r3g_bench$i.json").read(); d=json.loads(raw[raw.index("{"):])
pk=d["roofline"]["per_kernel"]
print("ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "bf16", round(d["bf16_mode"]["ms_per_step"],3), "padskip", round(d["pad_skip"]["ms_per_step"],3), d["pad_skip"]["valid_frames_bit_identical_to_headline_path"], "parity", d["parity_check"]["c2"]["modes"]["fp32"]["max_abs_mel_err_valid_frames"], d.get("errors"))
for k in ("ffn_fused","lfs2_attention_tc","predictor_pw_ln_gemm","qkv_gemm","out_proj_ln_gemm","lfs2_dwconv1d"): print("  ", k, pk[k]["ms"], pk[k]["frac"])
PY
done
