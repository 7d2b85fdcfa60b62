#!/usr/bin/env bash
# time window-attention library variants inside ONE gpurun call: bash profiles/run_window_ab.sh <tag> name1 name2 ...
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for n in "" "$@"; do
  lib=libwm_b200${n:+_$n}.so
  echo "== $lib" | tee -a gpurun_out/${tag}.txt
  WM_LIB_NAME=$lib timeout 300 python profiles/window_time.py 32 2>&1 | grep "^window" | tee -a gpurun_out/${tag}.txt
done
done
