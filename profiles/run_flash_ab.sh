#!/usr/bin/env bash
# flash-attention A/B inside ONE gpurun call: parity tests of every selectable generation, then per-launch times
# (batch 32; hd 64 + rel-pos, hd 128) alternating the generations.   usage: bash profiles/run_flash_ab.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash" -p no:cacheprovider > gpurun_out/${tag}_flash_tests.log 2>&1
echo "flash tests rc $?"; tail -5 gpurun_out/${tag}_flash_tests.log
for rep in 1 2; do
  for v in 4 7; do
    WM_FLASH_VERSION=$v timeout 300 python profiles/flash_time.py 2>&1 | tail -1 | tee -a gpurun_out/${tag}_flash_time.txt
  done
done
