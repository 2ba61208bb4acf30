set -x
mkdir -p gpurun_out
timeout 300 python tools/profile_train.py 3 fp32 C4 --table --world=8 > gpurun_out/r4g_train_c4_world8.txt 2>&1; echo "rc=$?"; head -50 gpurun_out/r4g_train_c4_world8.txt | grep -v Warn
