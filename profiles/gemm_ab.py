#!/usr/bin/env python
"""A/B of GEMM builds on three shapes: python profiles/gemm_ab.py  (WM_LIB_NAME selects the build)"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
M = 32 * 4096
def timeit(fn, iters=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
out = {}
for name, N, K, act, res in (("qkv", 2304, 768, 0, 0), ("lin1+gelu", 3072, 768, 1, 0), ("proj+res", 768, 768, 0, 1), ("lin2+res", 768, 3072, 0, 1)):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    o16 = None if res else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    x = torch.randn(M, N, device="cuda") if res else None
    for bn in (512, 256):
        out[f"{name}/{bn}"] = round(timeit(lambda: ops.gemm(a, w, bias, x, M if res else 0, o16, x, act, bn)), 4)
    ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out[f"{name}/cublas"] = round(timeit(lambda: torch.matmul(a, w.t(), out=ref)), 4)
print(os.environ.get("WM_LIB_NAME", "default"), json.dumps(out))
