set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "graphs" > gpurun_out/r2k_tests_graphs.log 2>&1; echo "graph tests rc=$?"
tail -5 gpurun_out/r2k_tests_graphs.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2k_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2k_bench_n1.err
