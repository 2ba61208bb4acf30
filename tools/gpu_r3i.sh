set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention_tc.py -q -m gpu -x > gpurun_out/r3i_tests_new.log 2>&1; echo "attn tests rc=$?"
tail -3 gpurun_out/r3i_tests_new.log
LFS2_ATTN_PP=2 timeout 900 python -m pytest tests/test_gpu_attention_tc.py -q -m gpu -x > gpurun_out/r3i_tests_pp2.log 2>&1; echo "attn tests pp2 rc=$?"
tail -3 gpurun_out/r3i_tests_pp2.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3i_tests_all.log 2>&1; echo "all tests rc=$?"
tail -3 gpurun_out/r3i_tests_all.log
timeout 300 python tools/profile_c3.py bf16 32 > gpurun_out/r3i_c3_profile_bf16.txt 2>&1; head -12 gpurun_out/r3i_c3_profile_bf16.txt
