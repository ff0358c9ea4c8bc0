#!/bin/bash
# Round-2 GPU pass H (8 GPUs): in-kernel peer-store exchange at full width -- frame identity + timing, strong-scaling bench.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 tools/frame_sharded_check.py > gpurun_out/h_frame_8gpu_peer.json 2> gpurun_out/h_frame_8gpu_peer.err
echo "exit $?" >> gpurun_out/h_frame_8gpu_peer.err
timeout 300 $TR --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/h_bench_8gpu_peer.json 2> gpurun_out/h_bench_8gpu_peer.err
echo "exit $?" >> gpurun_out/h_bench_8gpu_peer.err
tail -c 1200 gpurun_out/h_frame_8gpu_peer.json; tail -c 500 gpurun_out/h_frame_8gpu_peer.err; tail -c 300 gpurun_out/h_bench_8gpu_peer.err
