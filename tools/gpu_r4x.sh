set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r4x_bench_n4.json 2> gpurun_out/r4x_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4x_bench_n4.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'], 'n', d['n_gpus'])
t=d.get('train',{})
print('train',t.get('ms_per_step'),json.dumps(t.get('length_buckets',{}).get('runs'))[:300],json.dumps(t.get('bf16_mode'))[:200])
print('c3',d.get('c3_bf16',{}).get('value'))
print('errors',d.get('errors'))
PY
