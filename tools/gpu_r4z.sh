set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4z_smoke.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r4z_tests_all.log 2>&1; echo "all tests rc=$?"
tail -3 gpurun_out/r4z_tests_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --c3-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --parity-utts 0 --buckets --train-buckets 3 > gpurun_out/r4z_bench.json 2> gpurun_out/r4z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4z_bench.json') if l.startswith('{')][-1])
t=d.get('train',{})
print('value',d['value'],'train',t.get('ms_per_step'),t.get('gpu_launches_per_step'),json.dumps(t.get('length_buckets',{}).get('runs'))[:300], 'errors', d.get('errors'))
PY
