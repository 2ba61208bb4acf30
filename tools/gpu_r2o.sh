set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -x -k "multicast or two_pass_fp16 or test_linear" > gpurun_out/r2o_tests_pair.log 2>&1; echo "pair tests rc=$?"
tail -25 gpurun_out/r2o_tests_pair.log | cut -c1-250
timeout 300 python tools/gemm_ab.py knobs 2>&1 | tee gpurun_out/r2o_gemm_knobs.txt
