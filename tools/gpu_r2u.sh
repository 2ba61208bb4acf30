set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/parity_2gpu.py > gpurun_out/r2u_parity_2gpu.log 2> gpurun_out/r2u_parity_2gpu.err; echo "parity rc=$?"
cat gpurun_out/r2u_parity_2gpu.log | head -60
tail -5 gpurun_out/r2u_parity_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 --buckets > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2u_bench_n2.err
