set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_round2.py::test_c4_train_step_matches_oracle > gpurun_out/r2c_tests_all.log 2>&1; echo "all tests rc=$?"
tail -15 gpurun_out/r2c_tests_all.log
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -s > gpurun_out/r2c_tests_round2.log 2>&1; echo "round2 tests rc=$?"
grep -n "C4_P0\|max |mel\|resume\|passed\|failed\|Error" gpurun_out/r2c_tests_round2.log | head -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2c_bench_n1.err
python tools/attn_ab.py one default 2>&1 | tail -2
