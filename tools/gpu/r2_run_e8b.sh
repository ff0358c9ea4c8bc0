#!/bin/bash
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/e2_bench_8gpu.json 2> gpurun_out/e2_bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/e3_bench_8gpu.json 2> gpurun_out/e3_bench_8gpu.err
