mkdir -p gpurun_out
MOGP_LIB=$PWD/build/libmogp_trace.so timeout 300 python tools/chol_dtile_timeline.py > gpurun_out/s13_dtile.txt 2>&1
timeout 300 python tools/chol_i8_check.py c2 c3x4 c3 > gpurun_out/s13_chol.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s13_tests.log
tail -10 gpurun_out/s13_dtile.txt | cut -c1-200; grep "chol ms" gpurun_out/s13_chol.txt; cat gpurun_out/s13_tests.log
