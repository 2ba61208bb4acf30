set -x
mkdir -p gpurun_out
timeout 300 python tools/host_profile_train.py 8 > gpurun_out/r4e_host_profile_train.txt 2>&1; echo "rc=$?"; head -60 gpurun_out/r4e_host_profile_train.txt | grep -v Warn
