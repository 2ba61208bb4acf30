set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_forward.py tests/test_gpu_round2.py -q -m gpu -x > gpurun_out/r4q_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r4q_tests.log
timeout 300 python tools/profile_c3.py bf16 32 > gpurun_out/r4q_c3_profile_bf16.txt 2>&1; head -8 gpurun_out/r4q_c3_profile_bf16.txt | grep -v Warn
timeout 600 python bench.py --steps 10 --warmup 3 --train-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --buckets > gpurun_out/r4q_bench.json 2> gpurun_out/r4q_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r4q_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'])
c=d.get('c3_bf16',{})
print('c3',c.get('value'),c.get('ms_per_step'),json.dumps(c.get('pad_skip')))
print('parity c3', d.get('parity_check',{}).get('c3'))
print('errors', d.get('errors'))
PY
