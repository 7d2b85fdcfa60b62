#!/usr/bin/env bash
# ncu --set full on the qkv-shaped GEMM (CTA-pair kernel): DRAM traffic per launch for bench.py's roofline.traffic
mkdir -p gpurun_out
python profiles/gemm_one.py 131072 2304 768 512 0 0 3 > gpurun_out/ncu_plain_gemm_qkv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm -s 1 -c 1 -f -o gpurun_out/prof_gemm_qkv_r01z python profiles/gemm_one.py 131072 2304 768 512 0 0 3 > gpurun_out/ncu_gemm_qkv.log 2>&1
tail -1 gpurun_out/ncu_gemm_qkv.log
