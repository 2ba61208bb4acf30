set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sdp.py -q -m gpu -x -s > gpurun_out/r2v_tests_sdp.log 2>&1; echo "sdp tests rc=$?"
grep -n "sdp B=\|passed\|failed\|^E  " gpurun_out/r2v_tests_sdp.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2v_smoke.log
