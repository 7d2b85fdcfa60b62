#!/usr/bin/env bash
# time library variants inside ONE gpurun call: bash profiles/run_variants.sh <tag> <versions> name1 name2 ...
tag=$1; vers=$2; shift 2
mkdir -p gpurun_out
for rep in 1 2; do
for n in "" "$@"; do
  lib=libwm_b200${n:+_$n}.so
  echo "== $lib" | tee -a gpurun_out/${tag}.txt
  WM_LIB_NAME=$lib timeout 300 python profiles/flash_time2.py 32 $vers 2>&1 | grep "^v" | head -${NLINES:-1} | tee -a gpurun_out/${tag}.txt
done
done
