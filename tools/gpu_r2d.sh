set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hifigan.py tests/test_gpu_gemm_tc.py tests/test_gpu_attention_tc.py -q -m gpu -s > gpurun_out/r2d_tests_new.log 2>&1; echo "new tests rc=$?"
grep -n "hifigan \[\|attention t=\|passed\|failed\|^FAILED\|^E  " gpurun_out/r2d_tests_new.log | head -40
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2d_tests_all.log 2>&1; echo "all tests rc=$?"
tail -12 gpurun_out/r2d_tests_all.log
timeout 900 python bench.py --steps 10 --warmup 3 --train-steps 0 --c3-steps 0 --c1-steps 0 --c5-steps 0 --buckets > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2d_bench_n1.err
