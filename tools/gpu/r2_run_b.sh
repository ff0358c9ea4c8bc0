#!/bin/bash
# Round-2 GPU pass B: tests after the marcher (skipping, config table) and decoder (row windows, PDL, 2 streams) changes; timings.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu --maxfail=12 --durations=15 ) > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
python tools/sr_bench.py --no-ref --unit > gpurun_out/b_sr_bench.jsonl 2> gpurun_out/b_sr_bench.err
K4_SR_PDL=0 python tools/sr_bench.py --no-ref --unit >> gpurun_out/b_sr_bench.jsonl 2>> gpurun_out/b_sr_bench.err
python tools/quick_bench.py --hw 3024 4032 --modes ws --regimes fog shell --iters 3 > gpurun_out/b_quick_4k.jsonl 2>&1
K4_NO_SKIP=1 python tools/quick_bench.py --hw 3024 4032 --modes ws --regimes shell --iters 3 >> gpurun_out/b_quick_4k.jsonl 2>&1
python tools/quick_bench.py --kind cfgB --hw 3024 4032 --modes ws --regimes fog shell --iters 3 > gpurun_out/b_quick_mpi_4k.jsonl 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches_sr_tile520.csv \
    python tools/sr_one_tile.py > gpurun_out/b_sr_tile.log 2>&1
