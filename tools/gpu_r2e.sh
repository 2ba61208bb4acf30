set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_hifigan.py -q -m gpu -s > gpurun_out/r2e_tests_new.log 2>&1; echo "new tests rc=$?"
grep -n "wide attention\|hifigan \[\|passed\|failed\|^FAILED\|^E  " gpurun_out/r2e_tests_new.log | head -40
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2e_tests_all.log 2>&1; echo "all tests rc=$?"
tail -8 gpurun_out/r2e_tests_all.log
timeout 300 python tools/profile_c3.py bf16 32 > gpurun_out/r2e_c3_profile_bf16.txt 2>&1
timeout 300 python tools/profile_c3.py fp32 32 > gpurun_out/r2e_c3_profile_fp32.txt 2>&1
head -14 gpurun_out/r2e_c3_profile_bf16.txt; head -8 gpurun_out/r2e_c3_profile_fp32.txt
timeout 900 python bench.py --steps 10 --warmup 3 --train-steps 0 --c1-steps 0 --c5-steps 0 --buckets > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2e_bench_n1.err
