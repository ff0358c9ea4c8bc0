#!/bin/bash
# Round-2 GPU pass D: decoder after the SFT occupancy change (4 CTAs/SM): tests, frame timing, launch list.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sftnet.py tests/test_gpu_pipeline.py -q -m gpu > gpurun_out/d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest.log
python tools/sr_bench.py --no-ref --unit > gpurun_out/d_sr_bench.jsonl 2>&1
python tools/sr_bench.py --no-ref >> gpurun_out/d_sr_bench.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_sr_tile520.csv \
    python tools/sr_one_tile.py > gpurun_out/d_sr_tile.log 2>&1
