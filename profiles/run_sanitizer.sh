#!/usr/bin/env bash
# compute-sanitizer over the kernel unit tests at small shapes: memcheck, then racecheck on a subset.
mkdir -p gpurun_out
K='not benchmarked and not 131072 and not 262144 and not 32768 and not 40000 and not 20000 and not golden_bit_exact and not nms_edges'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "$K" -p no:cacheprovider > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc $?" | tee -a gpurun_out/r02_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r02_memcheck.log | tail -8
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm or transpose or patchify or add_cast or attn_small or postprocess or sigmoid_topk or nms_batched or hfc_finalize" -p no:cacheprovider > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck rc $?" | tee -a gpurun_out/r02_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r02_racecheck.log | tail -8
