#!/bin/bash
# 8 GPUs: the sharded frame with the decoder plan cached (the plan search was ~3 ms of Python per frame at 8 ranks)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29581 tools/frame_sharded_check.py > gpurun_out/h8b_frame_peer.json 2> gpurun_out/h8b_frame_peer.err
tail -c 1300 gpurun_out/h8b_frame_peer.json; grep -v "^\*\|OMP\|^$" gpurun_out/h8b_frame_peer.err | tail -4
