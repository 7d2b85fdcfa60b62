#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "window" > gpurun_out/r01o_ktests.log 2>&1
rc=$?; tail -12 gpurun_out/r01o_ktests.log
if [ $rc -ne 0 ]; then
  WM_LIB_NAME=libwm_b200_dbg.so CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "window" 2>&1 | grep -E "wm: mbarrier|FAILED|passed|failed|rror|assert" | sort | uniq -c | sort -rn | head -20
  exit 1
fi
NO_SHAPES=1 bash profiles/run_quick.sh r01o "window or flash"
