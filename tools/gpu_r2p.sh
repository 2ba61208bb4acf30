set -x
mkdir -p gpurun_out
QUICK="--steps 10 --warmup 3 --train-steps 0 --c3-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --parity-utts 0 --ref-utts 4 --ref-utts-max 4"
for v in 1 0; do
LFS2_GEMM_MULTICAST=$v timeout 600 python bench.py $QUICK > gpurun_out/r2p_bench_mc$v.json 2> gpurun_out/r2p_bench_mc$v.err; echo "bench rc=$?"
python - <<PY
import json
raw=open("gpurun_out/r2p_bench_mc$v.json").read(); d=json.loads(raw[raw.index("{"):])
pk=d["roofline"]["per_kernel"]
print("PAIR=$v ms/step", round(d["ms_per_step"],3), "bf16", round(d["bf16_mode"]["ms_per_step"],3))
for k in ("ffn_fused","qkv_gemm","out_proj_ln_gemm","predictor_pw_ln_gemm","mel_linear"): print("  ", k, pk[k]["ms"], pk[k]["frac"], pk[k]["frac_issued"])
PY
done
