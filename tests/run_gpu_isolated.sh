#!/usr/bin/env bash
# Diagnostic runner for the GPU box: each kernel-test group runs in its own process under `timeout`, so one
# trapped kernel (sticky CUDA error) or a hang cannot hide the results of the others.
# Usage (under gpurun): bash tests/run_gpu_isolated.sh [outfile]
out=${1:-gpurun_out/isolated.log}
mkdir -p "$(dirname "$out")"
: > "$out"
groups=("nms_batched" "gemm_matches" "gemm_epilogues or gemm_strided" "conv3x3" "layernorm or patchify or transpose or hfc_finalize or add_cast"
        "attn_small" "attn_flash_plain" "attn_flash_global" "attn_window" "postprocess or sigmoid_topk" "nms")
for g in "${groups[@]}"; do
  echo "=== group: $g" >> "$out"
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$g" -p no:cacheprovider 2>&1 | tail -40 >> "$out"
  echo "=== exit: ${PIPESTATUS[0]}" >> "$out"
done
grep -E "^=== |passed|failed|error" "$out"
