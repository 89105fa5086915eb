#!/bin/bash
# Build libmogp_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
SRC=mogp_emulator_b200/csrc
OUT=mogp_emulator_b200/libmogp_b200.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -Xptxas -v "$@" \
     $SRC/api.cu $SRC/chol.cu $SRC/kmat.cu $SRC/solve.cu $SRC/predict.cu $SRC/grad.cu $SRC/peak.cu $SRC/pool.cu $SRC/nccl_dyn.cu \
     -o $OUT -ldl 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "Wno" | head -20 || true
echo "built $OUT"
