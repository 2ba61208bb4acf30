set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_train.py -q -m gpu -s > gpurun_out/r4c_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r4c_tests.log; grep -h "train_length_buckets=\|bf16 train step" gpurun_out/r4c_tests.log
for nb in 1 3; do timeout 300 python tools/profile_train.py 3 fp32 C4 --table --buckets=$nb > gpurun_out/r4c_train_c4_b$nb.txt 2>&1; echo "b$nb rc=$?"; head -4 gpurun_out/r4c_train_c4_b$nb.txt | grep -v Warn; grep "attn_" gpurun_out/r4c_train_c4_b$nb.txt; done
LFS2_G2_OCC=1 timeout 300 python tools/profile_train.py 3 fp32 C4 --table --buckets=1 > gpurun_out/r4c_train_c4_b1_occ1.txt 2>&1; head -4 gpurun_out/r4c_train_c4_b1_occ1.txt | grep -v Warn; grep "attn_" gpurun_out/r4c_train_c4_b1_occ1.txt
