#!/bin/bash
# Round-2 GPU pass H (2 GPUs): the exchanges fused into the kernels (peer-mapped frames, NVLink stores) vs the all-gather path.
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/h_topo.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_marcher.py tests/test_gpu_sftnet.py -q -m gpu -x -k "frames or extra_destinations" > gpurun_out/h_pytest_1gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/h_pytest_1gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29541 tools/frame_sharded_check.py > gpurun_out/h_frame_2gpu_peer.json 2> gpurun_out/h_frame_2gpu_peer.err
echo "exit $?" >> gpurun_out/h_frame_2gpu_peer.err
K4_PEER=0 timeout 420 $TR --master-port 29542 tools/frame_sharded_check.py > gpurun_out/h_frame_2gpu_gather.json 2> gpurun_out/h_frame_2gpu_gather.err
echo "exit $?" >> gpurun_out/h_frame_2gpu_gather.err
timeout 420 $TR --master-port 29543 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/h_bench_2gpu_peer.json 2> gpurun_out/h_bench_2gpu_peer.err
echo "exit $?" >> gpurun_out/h_bench_2gpu_peer.err
K4_PEER=0 timeout 300 $TR --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/h_bench_2gpu_gather.json 2> gpurun_out/h_bench_2gpu_gather.err
echo "exit $?" >> gpurun_out/h_bench_2gpu_gather.err
tail -c 600 gpurun_out/h_pytest_1gpu.log; tail -c 1500 gpurun_out/h_frame_2gpu_peer.json; tail -c 800 gpurun_out/h_frame_2gpu_peer.err; tail -c 700 gpurun_out/h_frame_2gpu_gather.json
