set -x
mkdir -p gpurun_out
for v in 0 1 3; do LFS2_DWCONV_TMA=$v timeout 300 python tools/dwconv_ab.py > gpurun_out/r4m_dwconv_ab_$v.txt 2>&1; cat gpurun_out/r4m_dwconv_ab_$v.txt | grep -v Warn; done
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_train.py -q -m gpu -x > gpurun_out/r4m_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r4m_tests.log
