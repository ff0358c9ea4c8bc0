#!/bin/bash
# Round-2 GPU pass E (8 GPUs): sharded full frame identity + phase timing, contract bench at N=8.
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/frame_sharded_check.py > gpurun_out/e_frame_8gpu.json 2> gpurun_out/e_frame_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/e_bench_8gpu.json 2> gpurun_out/e_bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 tools/frame_sharded_check.py > gpurun_out/e_frame_4gpu.json 2> gpurun_out/e_frame_4gpu.err
