#!/bin/bash
# ncu evidence for the C3 bench step (run under gpurun; outputs land in gpurun_out/).
#   1. launch list with per-launch device time of one whole timed step
#   2. --set full captures of the dominant kernels
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -s 3075 -c 3080 --csv --log-file gpurun_out/launches_c3.csv $B > gpurun_out/launches_c3.log 2>&1
B0="python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:predict_trsm -c 1 -f -o gpurun_out/prof_trsm $B0 > gpurun_out/prof_trsm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chol_tile -s 20 -c 4 -f -o gpurun_out/prof_chol $B0 > gpurun_out/prof_chol.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:potf2|solve_alpha|kmat_kernel" -c 6 -f -o gpurun_out/prof_misc $B0 > gpurun_out/prof_misc.log 2>&1
ls -la gpurun_out
