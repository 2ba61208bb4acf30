set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_round2.py -q -m gpu > gpurun_out/r4y_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r4y_tests.log
timeout 300 python tools/profile_train.py 3 fp32 C4 --table > gpurun_out/r4y_train_c4.txt 2>&1; head -6 gpurun_out/r4y_train_c4.txt | grep -v Warn; grep "layernorm_bwd\|add_inplace" gpurun_out/r4y_train_c4.txt
