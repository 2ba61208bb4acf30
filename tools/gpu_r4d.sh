set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_round2.py -q -m gpu -x > gpurun_out/r4d_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r4d_tests.log
for nb in 1 3; do timeout 300 python tools/profile_train.py 3 fp32 C4 --table --buckets=$nb > gpurun_out/r4d_train_c4_b$nb.txt 2>&1; echo "b$nb rc=$?"; head -4 gpurun_out/r4d_train_c4_b$nb.txt | grep -v Warn; grep "attn_\|wgrad" gpurun_out/r4d_train_c4_b$nb.txt; done
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r4d_bench_n1.json 2> gpurun_out/r4d_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4d_bench_n1.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'])
print('c3',json.dumps(d.get('c3_bf16'))[:900])
t=d.get('train',{})
print('train',t.get('ms_per_step'),json.dumps(t.get('length_buckets'))[:700],json.dumps(t.get('bf16_mode'))[:500])
print('errors',d.get('errors'))
PY
