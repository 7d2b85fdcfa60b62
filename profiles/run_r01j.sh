#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" > gpurun_out/r01n_ktests.log 2>&1
rc=$?; tail -5 gpurun_out/r01n_ktests.log
if [ $rc -ne 0 ]; then
  WM_LIB_NAME=libwm_b200_dbg.so CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | grep -E "wm: mbarrier|FAILED|passed|failed|rror" | sort | uniq -c | sort -rn | head -20
  exit 1
fi
bash profiles/run_quick.sh r01n "gemm or conv"
