set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_train.py -q -m gpu > gpurun_out/r4t_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r4t_tests.log
