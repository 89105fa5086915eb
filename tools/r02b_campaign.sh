#!/bin/bash
# Second measurement campaign of round 2 on ONE B200 (run under gpurun), after the tcgen05 Cholesky and the diagonal-tile rework:
# GPU tests, the bench lines quoted in DESIGN.md, ncu evidence.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02b_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench_c3.err
timeout 600 python bench.py --workload c3a --no-other --no-cpu > gpurun_out/r02b_bench_c3_adaptive.json 2> gpurun_out/r02b_bench_c3_adaptive.err
timeout 600 python bench.py --workload c2 > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err
timeout 600 python bench.py --workload c4 --no-other > gpurun_out/r02b_bench_c4.json 2> gpurun_out/r02b_bench_c4.err
timeout 300 python tools/chol_i8_check.py c2 c3x4 c3 c4 > gpurun_out/r02b_chol_i8_check.txt 2>&1
B="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-other"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_c3.csv $B > gpurun_out/r02b_launches_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chol_i8_kernel -s 1 -c 1 -f -o gpurun_out/r02b_prof_chol_i8 $B > gpurun_out/r02b_prof_chol_i8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:i8_trsm_kernel -s 1 -c 1 -f -o gpurun_out/r02b_prof_i8_trsm $B > gpurun_out/r02b_prof_i8_trsm.log 2>&1
B4="python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu --no-e2e --no-other"
ncu --set full --clock-control none --import-source on -k regex:chol_i8_kernel -s 1 -c 1 -f -o gpurun_out/r02b_prof_chol_i8_c4 $B4 > gpurun_out/r02b_prof_chol_i8_c4.log 2>&1
ls -la gpurun_out | tail -20
