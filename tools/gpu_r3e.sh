set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu -x -k "synthesis_stream" > gpurun_out/r3e_tests_new.log 2>&1; echo "stream tests rc=$?"
tail -5 gpurun_out/r3e_tests_new.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3e_tests_all.log 2>&1; echo "all tests rc=$?"
tail -3 gpurun_out/r3e_tests_all.log
QUICK="--steps 10 --warmup 3 --train-steps 0 --c3-steps 0 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --parity-utts 8 --ref-utts 4 --ref-utts-max 4"
timeout 600 python bench.py $QUICK > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r3e_bench.err
python - <<'PY'
import json
raw=open("gpurun_out/r3e_bench.json").read(); d=json.loads(raw[raw.index("{"):])
print("ms/step", round(d["ms_per_step"],3), "value", d["value"], "e2e", json.dumps(d["e2e"]))
PY
