#!/usr/bin/env bash
# ncu --set full of the lin1 + GELU CTA-pair GEMM in isolation (profiles/gemm_gelu_time.py): bash profiles/run_gelu_ncu.sh <tag>
tag=${1:-r02y}
mkdir -p gpurun_out
timeout 200 python profiles/gemm_gelu_time.py > gpurun_out/${tag}_gelu_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_gelu_plain.log; exit 1; }
cat gpurun_out/${tag}_gelu_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 1 -c 1 -o gpurun_out/${tag}_prof_gelu python profiles/gemm_gelu_time.py > gpurun_out/${tag}_ncu_gelu.log 2>&1
tail -2 gpurun_out/${tag}_ncu_gelu.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/${tag}_prof_gelu.ncu-rep 12 > gpurun_out/${tag}_ncu_gelu_summary.txt 2>&1; head -16 gpurun_out/${tag}_ncu_gelu_summary.txt | cut -c1-400
python profiles/ncu_phases.py gpurun_out/${tag}_prof_gelu.ncu-rep 2>&1 | awk '$2+0 >= 0.4' | head -70
