set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fastdiff.py tests/test_gpu_length_regulator.py -q -m gpu -s > gpurun_out/r2j_tests_fastdiff.log 2>&1; echo "fastdiff tests rc=$?"
grep -n "fastdiff \[\|fastdiff teacher\|passed\|failed\|^FAILED\|^E  " gpurun_out/r2j_tests_fastdiff.log | head -40
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2j_tests_all.log 2>&1; echo "all tests rc=$?"
tail -6 gpurun_out/r2j_tests_all.log
