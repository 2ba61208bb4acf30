set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4u_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4u_smoke.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r4u_tests_all.log 2>&1; echo "all tests rc=$?"
tail -3 gpurun_out/r4u_tests_all.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r4u_bench_n1.json 2> gpurun_out/r4u_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4u_bench_n1.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for k in ('lfs2_attention_tc','ffn_fused','lfs2_dwconv1d'): print(k, d['roofline']['per_kernel'].get(k))
c=d.get('c3_bf16',{})
print('c3',c.get('value'),c.get('ms_per_step'),json.dumps(c.get('pad_skip')))
t=d.get('train',{})
print('train',t.get('ms_per_step'),json.dumps(t.get('length_buckets',{}).get('runs'))[:400])
print('errors',d.get('errors'))
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4u_bench_reference.json 2> gpurun_out/r4u_bench_reference.err; echo "ref rc=$?"
