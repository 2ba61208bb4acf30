set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_forward.py tests/test_gpu_round2.py tests/test_gpu_hifigan.py -q -m gpu -x > gpurun_out/r4p_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r4p_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --c3-steps 0 --train-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --buckets > gpurun_out/r4p_bench.json 2> gpurun_out/r4p_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r4p_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for k in ('lfs2_attention_tc','ffn_fused','predictor_pw_ln_gemm','out_proj_ln_gemm','qkv_gemm','lfs2_dwconv1d'): print(k, d['roofline']['per_kernel'].get(k))
print('parity', d.get('parity_check',{}).get('c2',{}).get('modes',{}).get('fp32'))
print('pad_skip', d.get('pad_skip',{}).get('ms_per_step'), 'bf16', d.get('bf16_mode',{}).get('ms_per_step'), 'errors', d.get('errors'))
PY
