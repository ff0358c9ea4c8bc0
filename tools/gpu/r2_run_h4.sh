#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29571 tools/frame_sharded_check.py --peer-breakdown > gpurun_out/h4_frame_breakdown.json 2> gpurun_out/h4_frame_breakdown.err
tail -c 2500 gpurun_out/h4_frame_breakdown.json; grep -v "^\*\|OMP\|^$" gpurun_out/h4_frame_breakdown.err | tail -5
