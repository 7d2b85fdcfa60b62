#!/usr/bin/env python
"""Run one wm_gemm_bf16 shape a few times (for ncu): python profiles/gemm_one.py M N K bn act res(0|1) iters"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
M, N, K, bn, act, res, iters = (int(x) for x in sys.argv[1:8])
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
o16 = None if res else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
x = torch.randn(M, N, device="cuda") if res else None
for _ in range(iters):
    ops.gemm(a, w, bias, x, M if res else 0, o16, x, act, bn)
torch.cuda.synchronize()
print("ok")
