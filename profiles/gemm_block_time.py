#!/usr/bin/env python
"""The four GEMMs of a ViT-B encoder block at batch 32 (CTA-pair kernel), for in-run A/B of differently built libraries."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
M = 131072
out = []
for name, N, K, act, res in (("qkv", 2304, 768, 0, 0), ("proj+res", 768, 768, 0, 1), ("lin1+gelu", 3072, 768, 1, 0), ("lin2+res", 768, 3072, 0, 1)):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    o16 = None if res else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    x = torch.randn(M, N, device="cuda") if res else None
    f = lambda: ops.gemm(a, w, bias, x, M if res else 0, o16, x, act, 512)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        f()
    e.record(); torch.cuda.synchronize()
    out.append(f"{name} {s.elapsed_time(e) / 10:.4f}")
    del a, w, bias, o16, x
print(os.environ.get("WM_LIB_NAME", "libwm_b200.so"), " | ".join(out))
