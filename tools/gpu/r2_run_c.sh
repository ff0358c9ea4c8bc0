#!/bin/bash
# Round-2 GPU pass C (2 GPUs): full GPU tests, sharded-frame identity + phase timing at N=2, contract bench at N=1 and N=2.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/scale_parity.jsonl
( time python -m pytest tests -q -m gpu --maxfail=12 --durations=10 ) > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/frame_sharded_check.py > gpurun_out/c_frame_2gpu.json 2> gpurun_out/c_frame_2gpu.err
python tools/frame_sharded_check.py > gpurun_out/c_frame_1gpu.json 2> gpurun_out/c_frame_1gpu.err
python bench.py --steps 10 --warmup 3 > gpurun_out/c_bench_1gpu.json 2> gpurun_out/c_bench_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c_bench_2gpu.json 2> gpurun_out/c_bench_2gpu.err
python tools/sr_bench.py --no-ref > gpurun_out/c_sr_bench.jsonl 2>&1
