#!/usr/bin/env bash
# whole-step A/B of library variants inside ONE gpurun call: bash profiles/run_bench_ab.sh <tag> name1 name2 ...   ("" = default lib first)
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2 3; do
for n in "" "$@"; do
  lib=libwm_b200${n:+_$n}.so
  WM_LIB_NAME=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); b = d['breakdown_ms_per_step']
print('$lib', round(d['value'], 1), round(d['ms_per_step'], 3), 'ln', b.get('layernorm'), 'gemm', b.get('gemm'), 'flash', b.get('attn_flash'), d['clocks']['sm_mhz'])" | tee -a gpurun_out/${tag}.txt
done
done
