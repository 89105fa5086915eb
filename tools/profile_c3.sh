#!/bin/bash
# ncu evidence for the C3 bench step (run under gpurun; outputs land in gpurun_out/).
#   1. launch list with per-launch device time of one warm-up + one timed step
#   2. --set full captures of the dominant kernels (second instance = the timed step's)
set -x
mkdir -p gpurun_out
W=${1:-c3}
B="python bench.py --workload $W --steps 1 --warmup 1 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$W.csv $B > gpurun_out/launches_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:predict_trsm -s 1 -c 1 -f -o gpurun_out/prof_trsm_$W $B > gpurun_out/prof_trsm_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chol_dataflow -s 1 -c 1 -f -o gpurun_out/prof_chol_$W $B > gpurun_out/prof_chol_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:solve_alpha|kmat_kernel|mean_reduce" -s 4 -c 4 -f -o gpurun_out/prof_misc_$W $B > gpurun_out/prof_misc_$W.log 2>&1
ls -la gpurun_out
