#!/bin/bash
# ncu evidence for the C3 bench step on the int8 tcgen05 path (run under gpurun; outputs land in gpurun_out/).
#   1. launch list with per-launch device time of one warm-up + one timed step
#   2. --set full captures of i8_row_kernel of the timed step: block row 16 (the average launch) and 31 (the largest)
set -x
mkdir -p gpurun_out
B="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3_i8.csv $B > gpurun_out/launches_c3_i8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:i8_row_kernel -s 48 -c 1 -f -o gpurun_out/prof_i8_row16 $B > gpurun_out/prof_i8_row16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:i8_row_kernel -s 63 -c 1 -f -o gpurun_out/prof_i8_row31 $B > gpurun_out/prof_i8_row31.log 2>&1
ls -la gpurun_out
