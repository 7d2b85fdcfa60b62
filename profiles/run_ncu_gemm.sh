#!/usr/bin/env bash
# ncu --set full on the qkv-shaped GEMM: single-CTA (bn 256) and CTA-pair (bn 512) kernels, and lin1+GELU
mkdir -p gpurun_out
for cfg in "2304 768 256 0 0 q256" "2304 768 512 0 0 q512" "3072 768 512 1 0 g512" "768 768 512 0 1 r512"; do
  set -- $cfg
  python profiles/gemm_one.py 131072 $1 $2 $3 $4 $5 3 > gpurun_out/ncu_plain_gemm_$6.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gemm -s 1 -c 1 -o gpurun_out/prof_gemm_$6 python profiles/gemm_one.py 131072 $1 $2 $3 $4 $5 3 > gpurun_out/ncu_gemm_$6.log 2>&1
  tail -1 gpurun_out/ncu_gemm_$6.log
done
