#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash" > gpurun_out/r01g_ktests.log 2>&1 || { tail -30 gpurun_out/r01g_ktests.log; WM_LIB_NAME=libwm_b200_dbg.so CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash" 2>&1 | grep -E "wm: mbarrier|FAILED|passed|failed" | sort | uniq -c | head; exit 1; }
tail -1 gpurun_out/r01g_ktests.log
for turns in 1 0; do
  WM_FLASH_TURNS=$turns WM_LIB_NAME=libwm_b200_dbg.so python profiles/flash_trace.py 64 1 > gpurun_out/flash_trace_64_t$turns.txt 2>&1; echo "== turns=$turns hd64"; tail -5 gpurun_out/flash_trace_64_t$turns.txt
  WM_FLASH_TURNS=$turns WM_LIB_NAME=libwm_b200_dbg.so python profiles/flash_trace.py 128 0 > gpurun_out/flash_trace_128_t$turns.txt 2>&1; echo "== turns=$turns hd128"; tail -4 gpurun_out/flash_trace_128_t$turns.txt
  WM_FLASH_TURNS=$turns timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01g_t$turns.json 2> gpurun_out/bench_r01g_t$turns.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r01g_t$turns.json')); print('turns=$turns', d['value'], d['e2e']['value']); print(d['breakdown_ms_per_step'])"
done
