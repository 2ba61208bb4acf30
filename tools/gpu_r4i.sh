set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r4i_tests_all.log 2>&1; echo "all tests rc=$?"
tail -4 gpurun_out/r4i_tests_all.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r4i_bench_n1.json 2> gpurun_out/r4i_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4i_bench_n1.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'])
c=d.get('c3_bf16',{})
print('c3',c.get('value'),c.get('ms_per_step'),json.dumps(c.get('bucketed'))[:400],json.dumps(c.get('pad_skip')))
t=d.get('train',{})
print('train',t.get('ms_per_step'),json.dumps(t.get('length_buckets',{}).get('runs'))[:500])
print('errors',d.get('errors'))
PY
