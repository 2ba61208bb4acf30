set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_train.py -q -m gpu -x > gpurun_out/r4t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r4t_tests.log
timeout 300 python tools/profile_train.py 3 fp32 C4 --table --world=8 > gpurun_out/r4t_train_c4_world8.txt 2>&1; head -8 gpurun_out/r4t_train_c4_world8.txt | grep -v Warn; grep "weight_prep\|transpose\|split_bf16" gpurun_out/r4t_train_c4_world8.txt
timeout 300 python tools/profile_train.py 3 fp32 C4 --table > gpurun_out/r4t_train_c4.txt 2>&1; head -4 gpurun_out/r4t_train_c4.txt | grep -v Warn; grep "weight_prep\|transpose\|split_bf16" gpurun_out/r4t_train_c4.txt
