set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --c3-steps 0 --train-steps 0 --c1-steps 0 --c5-steps 0 --parity-utts 0 --buckets > gpurun_out/r4w_bench.json 2> gpurun_out/r4w_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r4w_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'])
v=d.get('vocoder_hifigan',{})
print('vocoder', v.get('value'), v.get('ms_per_batch'), v.get('bf16_mode'), json.dumps(v.get('stream_mel_and_wav')))
print('errors', d.get('errors'))
PY
