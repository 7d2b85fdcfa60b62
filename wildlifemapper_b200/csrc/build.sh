#!/usr/bin/env bash
# Build libwm_b200.so for sm_100a (in-tree; the .so is git-ignored but travels to the GPU box).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall ${WM_NVCC_EXTRA:-}"
BUILD=${WM_BUILD_DIR:-build}
LIBNAME=${WM_LIB_NAME:-libwm_b200.so}
mkdir -p $BUILD
pids=()
SRCS="gemm attn_flash4 attn_flash7 attn_window2 attn_window3 elementwise postprocess frontend criterion api"
for f in $SRCS; do
  if [ ! -f $BUILD/$f.o ] || [ $f.cu -nt $BUILD/$f.o ] || [ common.cuh -nt $BUILD/$f.o ] || [ wm_internal.h -nt $BUILD/$f.o ] || [ ../../include/wm_b200.h -nt $BUILD/$f.o ]; then
    $NVCC $FLAGS -Xptxas -v -c $f.cu -o $BUILD/$f.o > $BUILD/$f.log 2>&1 &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [ -z "$p" ] || wait $p || rc=1; done
if [ $rc -ne 0 ]; then cat $BUILD/*.log | grep -E "error|Error" -B2 -A6 | head -80; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../$LIBNAME $(for f in $SRCS; do echo $BUILD/$f.o; done) -lcudart
echo "built $(cd .. && pwd)/$LIBNAME"
