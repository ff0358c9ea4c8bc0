#!/bin/bash
# 2 GPUs: is the slower decode of the FIRST regime in peer mode tied to lazy peer-access enabling or to the data?
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29561 tools/frame_sharded_check.py > gpurun_out/h2b_eager_shell_first.json 2> gpurun_out/h2b_eager_shell_first.err
K4_PEER_EAGER=0 timeout 300 $TR --master-port 29562 tools/frame_sharded_check.py --fog-first > gpurun_out/h2b_lazy_fog_first.json 2> gpurun_out/h2b_lazy_fog_first.err
tail -c 900 gpurun_out/h2b_eager_shell_first.json; tail -c 900 gpurun_out/h2b_lazy_fog_first.json; tail -c 300 gpurun_out/h2b_eager_shell_first.err
