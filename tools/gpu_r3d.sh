set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --c1-steps 0 --c5-steps 0 --vocoder-utts 0 > gpurun_out/r3d_bench_n8.json 2> gpurun_out/r3d_bench_n8.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r3d_bench_n8.err
