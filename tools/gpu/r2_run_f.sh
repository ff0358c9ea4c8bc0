#!/bin/bash
# Round-2 GPU pass F: full tests after the centre-out tile order, contract bench (N=1) with the ncu traffic figure, decoder streams sweep.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/scale_parity.jsonl
( time python -m pytest tests -q -m gpu --maxfail=12 --durations=8 ) > gpurun_out/f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench_1gpu.err
python tools/sr_bench.py --no-ref > gpurun_out/f_sr_bench.jsonl 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
