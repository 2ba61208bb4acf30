set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_round2.py -q -m gpu -x > gpurun_out/r4s_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r4s_tests.log
timeout 300 python tools/profile_train.py 3 fp32 C4 --table --world=8 > gpurun_out/r4s_train_c4_world8.txt 2>&1; head -8 gpurun_out/r4s_train_c4_world8.txt | grep -v Warn
timeout 300 python tools/profile_train.py 3 fp32 C4 --table > gpurun_out/r4s_train_c4.txt 2>&1; head -6 gpurun_out/r4s_train_c4.txt | grep -v Warn; grep layernorm_bwd gpurun_out/r4s_train_c4.txt
timeout 300 python tools/profile_train.py 3 fp32 C4 --table --buckets=3 > gpurun_out/r4s_train_c4_b3.txt 2>&1; head -4 gpurun_out/r4s_train_c4_b3.txt | grep -v Warn; grep layernorm_bwd gpurun_out/r4s_train_c4_b3.txt
