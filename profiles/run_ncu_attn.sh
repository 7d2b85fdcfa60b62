#!/usr/bin/env bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"flash3_kernel|window2_kernel" -s 1 -c 2 -o gpurun_out/prof_attn_r01o $CMD > gpurun_out/ncu_attn_r01o.log 2>&1
tail -2 gpurun_out/ncu_attn_r01o.log
