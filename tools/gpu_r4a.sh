set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_train.py -q -m gpu -x -s > gpurun_out/r4a_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r4a_tests.log; grep -h "train_length_buckets=\|C4_P0 train step" gpurun_out/r4a_tests.log
timeout 300 python tools/train_bf16_check.py > gpurun_out/r4a_bf16_check.txt 2>&1; echo "bf16 check rc=$?"; grep "whole gradient\|median" gpurun_out/r4a_bf16_check.txt
for nb in 1 2 3 4; do timeout 300 python tools/profile_train.py 3 fp32 C4 --table --buckets=$nb > gpurun_out/r4a_train_c4_b$nb.txt 2>&1; echo "b$nb rc=$?"; head -12 gpurun_out/r4a_train_c4_b$nb.txt | grep -v Warn; done
timeout 300 python tools/profile_train.py 3 bf16 C4 --table --buckets=1 > gpurun_out/r4a_train_c4_bf16.txt 2>&1; head -3 gpurun_out/r4a_train_c4_bf16.txt
