set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_abi_cpu.py -x -q -m gpu -s > gpurun_out/r2b_tests_new.log 2>&1; echo "new tests rc=$?"
tail -5 gpurun_out/r2b_tests_new.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_bench_n1.err
timeout 300 python tools/profile_c3.py bf16 32 > gpurun_out/r2b_c3_profile.txt 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_tests_all.log 2>&1; echo "all tests rc=$?"
tail -3 gpurun_out/r2b_tests_all.log
