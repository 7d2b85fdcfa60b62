#!/usr/bin/env bash
# event traces of one CTA of the hd-64 + rel-pos flash kernels (diagnostics build libwm_b200_trace.so, -DWM_F3_TRACE)
mkdir -p gpurun_out
for v in 4 7; do
  echo "=== flash_version $v" | tee -a gpurun_out/${1:-r02}_flash_trace.txt
  WM_LIB_NAME=libwm_b200_trace.so WM_FLASH_VERSION=$v timeout 300 python profiles/flash4_trace.py 64 1 2>&1 | tail -14 | tee -a gpurun_out/${1:-r02}_flash_trace.txt
done
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash" -p no:cacheprovider 2>&1 | tail -3
