set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -q -m gpu -x > gpurun_out/r4k_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r4k_tests.log
for v in 1 0; do
LFS2_DWCONV_TMA=$v timeout 600 python bench.py --steps 10 --warmup 3 --c3-steps 0 --train-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --parity-utts 0 --buckets > gpurun_out/r4k_bench_tma$v.json 2> gpurun_out/r4k_bench_tma$v.err; echo "bench tma=$v rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r4k_bench_tma$v.json') if l.startswith('{')][-1])
print('tma=$v value',d['value'],'ms',d['ms_per_step'],'dwconv',d['roofline']['per_kernel'].get('lfs2_dwconv1d'), 'pad_skip', d.get('pad_skip',{}).get('ms_per_step'), 'errors', d.get('errors'))
PY
done
timeout 300 python tools/profile_c3.py bf16 32 > gpurun_out/r4k_c3_profile_bf16.txt 2>&1; head -12 gpurun_out/r4k_c3_profile_bf16.txt | grep -v Warn
