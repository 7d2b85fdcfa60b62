#!/usr/bin/env bash
# ncu --set full on lin1+GELU (CTA-pair kernel) next to the same shape without activation
mkdir -p gpurun_out
for cfg in "3072 768 512 1 0 gelu" "3072 768 512 0 0 noact"; do
  set -- $cfg
  python profiles/gemm_one.py 131072 $1 $2 $3 $4 $5 3 > gpurun_out/ncu_plain_gemm_$6.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gemm -s 1 -c 1 -f -o gpurun_out/prof_gemm_$6 python profiles/gemm_one.py 131072 $1 $2 $3 $4 $5 3 > gpurun_out/ncu_gemm_$6.log 2>&1
  tail -1 gpurun_out/ncu_gemm_$6.log
done
