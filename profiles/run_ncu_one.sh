#!/usr/bin/env bash
# Usage (under gpurun): bash profiles/run_ncu_one.sh <tag> <kernel-regex> [skip] [count]
tag=$1; rx=$2; skip=${3:-1}; cnt=${4:-2}
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/ncu_plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
