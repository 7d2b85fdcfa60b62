#!/usr/bin/env python
"""Per-shape timing of wm_gemm_bf16 on the GEMM shapes of the ViT-B batch-32 step (M = 32*4096), next to
torch.matmul (cuBLAS, bf16 out, no epilogue) on the same operands.  CUDA events, L2 flushed by operand size.
Usage (GPU box): python profiles/gemm_shapes.py [batch] > gpurun_out/gemm_shapes.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M = B * 4096
dev = "cuda"
# name, M, N, K, bias, residual(fp32, rows), out_bf16, out_f32, act
SHAPES = [
    ("qkv", M, 2304, 768, 1, 0, 1, 0, 0),
    ("proj+res", M, 768, 768, 1, M, 0, 1, 0),
    ("lin1+gelu", M, 3072, 768, 1, 0, 1, 0, 1),
    ("lin1 (no act)", M, 3072, 768, 1, 0, 1, 0, 0),
    ("lin2+res", M, 768, 3072, 1, M, 0, 1, 0),
    ("patch+pos", M, 768, 768, 1, 4096, 1, 1, 0),
    ("hfc_embed", M, 1024, 256, 1, 0, 1, 0, 0),
    ("hfc 1024x1024", M, 1024, 1024, 1, 0, 1, 0, 0),
    ("hfc kv", M, 2048, 1024, 1, 0, 1, 0, 0),
    ("hfc out+res", M, 1024, 1024, 1, M, 0, 1, 0),
    ("hfc lin1+relu", M, 1024, 1024, 1, 0, 1, 0, 2),
    ("proj_patch dual", M, 1024, 768, 1, 0, 1, 1, 0),
    ("proj_back+res", M, 768, 1024, 1, M, 0, 1, 0),
    ("lowpass1", B * 1024, 2048, 1024, 0, 0, 1, 0, 0),
    ("lowpass2", B * 1024, 1024, 2048, 0, 0, 0, 1, 0),
    ("neck0", M, 256, 768, 0, 0, 0, 1, 0),
]


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


rows = []
for name, m, n, k, has_bias, res_rows, o16, o32, act in SHAPES:
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    w = (torch.randn(n, k, device=dev) * k ** -0.5).to(torch.bfloat16)
    bias = torch.randn(n, device=dev) if has_bias else None
    res = torch.randn(res_rows, n, device=dev) if res_rows else None
    out16 = torch.empty(m, n, device=dev, dtype=torch.bfloat16) if o16 else None
    out32 = torch.empty(m, n, device=dev) if o32 else None
    inplace = res_rows == m  # the residual stream is updated in place in the engine
    if inplace:
        out32 = res
    flop = 2.0 * m * n * k
    r = {"name": name, "M": m, "N": n, "K": k}
    for bn in (512, 256):
        if n < min(bn, 256) or (bn == 512 and n % 256):
            continue
        ms = timeit(lambda: ops.gemm(a, w, bias, res, res_rows, out16, out32, act, bn))
        r[f"wm_bn{bn}_ms"] = round(ms, 4)
        r[f"wm_bn{bn}_tflops"] = round(flop / ms / 1e9, 1)
    ref_out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: torch.matmul(a, w.t(), out=ref_out))
    r["cublas_ms"], r["cublas_tflops"] = round(ms, 4), round(flop / ms / 1e9, 1)
    byts = m * k * 2 + n * k * 2 + (m * n * 2 if o16 else 0) + (m * n * 4 if o32 else 0) + (m * n * 4 if inplace else 0)
    r["min_hbm_ms"] = round(byts / 6550.7e9 * 1e3, 4)
    rows.append(r)
    print(json.dumps(r), flush=True)
    del a, w, bias, res, out16, out32, ref_out
