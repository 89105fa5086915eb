#!/bin/bash
# Build libmogp_b200.so (sm_100a only) in-tree: one nvcc -c per source in parallel, then link.
# Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
SRC=mogp_emulator_b200/csrc
OUT=${OUT:-mogp_emulator_b200/libmogp_b200.so}      # OBJ=build/obj_trace OUT=build/libmogp_trace.so ./build.sh -DCHOL_TRACE: a debug build beside the product
OBJ=${OBJ:-build/obj}
mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
pids=()
for f in $SRC/*.cu; do
    b=$(basename $f .cu)
    if [ ! -f $OBJ/$b.o ] || [ -n "$(find $f $SRC/*.h $SRC/*.cuh include/mogp_b200.h -newer $OBJ/$b.o 2>/dev/null)" ] || [ -n "$*" ]; then
        ( nvcc $FLAGS "$@" -c $f -o $OBJ/$b.o > $OBJ/$b.log 2>&1 || { cat $OBJ/$b.log; exit 1; } ) &
        pids+=($!)
    fi
done
for p in "${pids[@]}"; do wait $p; done
cat $OBJ/*.log > build.log 2>/dev/null || true
nvcc -gencode arch=compute_100a,code=sm_100a -shared $OBJ/*.o -o $OUT -ldl
grep -E "error|warning" build.log | grep -v "Wno" | head -20 || true
echo "built $OUT"
