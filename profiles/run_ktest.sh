#!/usr/bin/env bash
# usage: bash profiles/run_ktest.sh <tag> <pytest -k expr>: kernel tests; on failure rerun with the diagnostics build
tag=$1; kexpr=$2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "$kexpr" > gpurun_out/${tag}_ktests.log 2>&1
rc=$?; tail -15 gpurun_out/${tag}_ktests.log
if [ $rc -ne 0 ]; then
  WM_LIB_NAME=libwm_b200_dbg.so CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "$kexpr" 2>&1 | grep -E "wm: mbarrier|FAILED|passed|failed|rror|assert" | sort | uniq -c | sort -rn | head -20
fi
exit $rc
