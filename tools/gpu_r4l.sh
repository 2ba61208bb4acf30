set -x
mkdir -p gpurun_out
for v in 0 1 2; do LFS2_DWCONV_TMA=$v timeout 300 python tools/dwconv_ab.py > gpurun_out/r4l_dwconv_ab_$v.txt 2>&1; cat gpurun_out/r4l_dwconv_ab_$v.txt | grep -v Warn; done
LFS2_DWCONV_TMA=2 timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k dwconv > gpurun_out/r4l_tests2.log 2>&1; echo "tests(2) rc=$?"; tail -2 gpurun_out/r4l_tests2.log
