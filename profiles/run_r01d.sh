#!/usr/bin/env bash
# flash v3 validation: kernel tests, then model tests + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash" > gpurun_out/r01d_flash_tests.log 2>&1
rc=$?
tail -15 gpurun_out/r01d_flash_tests.log
if [ $rc -ne 0 ]; then
  echo "== flash tests failed; rerunning with the debug library"
  WM_LIB_NAME=libwm_b200_dbg.so CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash" > gpurun_out/r01d_flash_tests_dbg.log 2>&1
  grep -E "wm: mbarrier|FAILED|passed|failed|Error|error" gpurun_out/r01d_flash_tests_dbg.log | sort | uniq -c | sort -rn | head -40
  exit 1
fi
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r01d_gpu_tests.log 2>&1; tail -4 gpurun_out/r01d_gpu_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_r01d.json > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r01d.json')); print(d['value'], d['e2e']['value'], d['breakdown_ms_per_step'])"; tail -3 gpurun_out/bench_r01d.err
