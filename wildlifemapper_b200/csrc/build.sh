#!/usr/bin/env bash
# Build libwm_b200.so for sm_100a (in-tree; the .so is git-ignored but travels to the GPU box).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall ${WM_NVCC_EXTRA:-}"
mkdir -p build
pids=()
for f in gemm attn_flash attn_flash2 attn_window elementwise postprocess api; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ wm_internal.h -nt build/$f.o ] || [ ../../include/wm_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS -Xptxas -v -c $f.cu -o build/$f.o > build/$f.log 2>&1 &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [ -z "$p" ] || wait $p || rc=1; done
if [ $rc -ne 0 ]; then cat build/*.log | grep -E "error|Error" -B2 -A6 | head -80; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libwm_b200.so build/gemm.o build/attn_flash.o build/attn_flash2.o build/attn_window.o build/elementwise.o build/postprocess.o build/api.o -lcudart
echo "built $(cd .. && pwd)/libwm_b200.so"
