#!/bin/bash
# Round-2 GPU pass G: SFT layers fused into the convolution epilogues: layout probe, decoder tests, timing vs unfused.
set -u
mkdir -p gpurun_out
./tools/micro/tmem_ld_shapes.bin > gpurun_out/g_tmem_probe.json 2>&1
timeout 600 python -m pytest tests/test_gpu_sftnet.py tests/test_gpu_pipeline.py -q -m gpu -x > gpurun_out/g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/g_pytest.log
timeout 300 python tools/sr_bench.py --no-ref --unit > gpurun_out/g_sr_bench.jsonl 2>&1
K4_SR_FUSE_SFT=0 timeout 300 python tools/sr_bench.py --no-ref --unit >> gpurun_out/g_sr_bench.jsonl 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_launches_sr_tile520.csv \
    python tools/sr_one_tile.py > gpurun_out/g_sr_tile.log 2>&1
