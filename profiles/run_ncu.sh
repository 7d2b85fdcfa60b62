#!/usr/bin/env bash
# Run under gpurun (1 GPU).  Usage: bash profiles/run_ncu.sh <tag>
# 1) plain run (must exit 0), 2) launch list with per-launch device time, 3) --set full capture of the top kernels.
tag=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/ncu_plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 30 -c 3 -o gpurun_out/prof_gemm_$tag $CMD > gpurun_out/ncu_gemm_$tag.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flash_attn_kernel -s 1 -c 2 -o gpurun_out/prof_flash_$tag $CMD > gpurun_out/ncu_flash_$tag.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:window_attn_kernel -s 1 -c 1 -o gpurun_out/prof_window_$tag $CMD > gpurun_out/ncu_window_$tag.log 2>&1
ls -la gpurun_out/
