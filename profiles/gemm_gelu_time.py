#!/usr/bin/env python
"""lin1 + GELU (M = 131072, N = 3072, K = 768, CTA-pair kernel) launch time, for in-run A/B of differently built libraries."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
M, N, K = 131072, 3072, 768
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
out = []
for act in (1, 0):
    for _ in range(3):
        ops.gemm(a, w, bias, None, 0, o, None, act, 512)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.gemm(a, w, bias, None, 0, o, None, act, 512)
    e.record(); torch.cuda.synchronize()
    out.append(f"act={act}: {s.elapsed_time(e) / 10:.4f} ms")
print(os.environ.get("WM_LIB_NAME", "libwm_b200.so"), " | ".join(out))
