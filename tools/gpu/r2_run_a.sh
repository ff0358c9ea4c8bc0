#!/bin/bash
# Round-2 GPU pass A: full GPU test suite (incl. the BASELINE-scale parity tests), contract bench, decoder launch list
# + ncu --set full of one residual dense block and of the high-resolution tail, DRAM traffic of the 4K marcher launch.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
( time python -m pytest tests -q -m gpu --maxfail=10 --durations=25 ) > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench exit $?" >> gpurun_out/a_bench.err
# decoder launch list (second forward of two)
ncu --metrics gpu__time_duration.sum --clock-control none -s 127 -c 127 --csv --log-file gpurun_out/a_launches_sr_tile520.csv \
    python tools/sr_one_tile.py > gpurun_out/a_sr_tile.log 2>&1
# one residual dense block of the second forward: sft<64>, conv<32> x4, sft<32>, conv<64>
ncu --set full --clock-control none --import-source on -k regex:'conv3x3|sft_tc' -s 125 -c 7 -f -o gpurun_out/a_rdb \
    python tools/sr_one_tile.py > gpurun_out/a_ncu_rdb.log 2>&1
# the tail of the second forward: sftbody, conv_body, up1 x4, up2 x4, conv_hr, conv_last
ncu --set full --clock-control none --import-source on -k regex:'conv3x3|sft_tc' -s 235 -c 13 -f -o gpurun_out/a_tail \
    python tools/sr_one_tile.py > gpurun_out/a_ncu_tail.log 2>&1
# DRAM traffic + binding-resource percentages of the 4K FOG marcher launch (the bench's launch)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum \
    --clock-control none -k regex:k4_march_ws -s 1 -c 1 --csv --log-file gpurun_out/a_march4k_metrics.csv \
    python tools/quick_bench.py --hw 3024 4032 --modes ws --regimes fog --iters 1 > gpurun_out/a_march4k.log 2>&1
ls -la gpurun_out > gpurun_out/a_ls.txt
