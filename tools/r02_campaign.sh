#!/bin/bash
# Round-2 measurement campaign on ONE B200 (run under gpurun): GPU tests, the bench lines quoted in DESIGN.md, ncu evidence.
# Outputs land in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
MOGP_TRSM_I8=0 timeout 600 python bench.py --no-other --no-cpu > gpurun_out/r02_bench_c3_fp64.json 2> gpurun_out/r02_bench_c3_fp64.err
timeout 600 python bench.py --workload c3a --no-other --no-cpu > gpurun_out/r02_bench_c3_adaptive.json 2> gpurun_out/r02_bench_c3_adaptive.err
MOGP_I8_CHECK=0 timeout 600 python bench.py --no-other --no-cpu --no-e2e > gpurun_out/r02_bench_c3_nocheck.json 2> gpurun_out/r02_bench_c3_nocheck.err
timeout 600 python bench.py --workload c2 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
timeout 600 python bench.py --workload c4 --no-other > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
B="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-other"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3.csv $B > gpurun_out/r02_launches_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:i8_trsm_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_i8_trsm_final $B > gpurun_out/r02_prof_i8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chol_dataflow -s 1 -c 1 -f -o gpurun_out/r02_prof_chol $B > gpurun_out/r02_prof_chol.log 2>&1
ncu --set full --clock-control none -k "regex:kmat_kernel|i8_slice|solve_alpha|predict_trsm|mean_reduce|i8_check" -s 6 -c 8 -f -o gpurun_out/r02_prof_misc $B > gpurun_out/r02_prof_misc.log 2>&1
bash tools/sanitize.sh > gpurun_out/r02_sanitizer.log 2>&1
ls -la gpurun_out | tail -30
