#!/usr/bin/env python
"""Launch time of the global-attention kernel (ViT-B: 12 heads x 64, 4096 tokens, batch 32, rel-pos) and of the HFC
cross-attention shape (8 heads x 128), for A/B runs of differently built libraries inside ONE gpurun call:
  WM_LIB_NAME=libwm_b200_a.so python profiles/flash_time.py"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
B, T = 32, 4096
res = []
for H, hd, relpos in ((12, 64, 1), (8, 128, 0)):
    D = H * hd
    qkv = (torch.randn(B * T, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
    out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
    table = (torch.randn(256, hd, device="cuda") * 0.05).to(torch.bfloat16) if relpos else None
    f = lambda: ops.attn_flash(qkv, 0, qkv, D, qkv, 2 * D, table, out, B, H, T, T, hd, 1 / math.sqrt(hd))
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        f()
    e.record(); torch.cuda.synchronize()
    res.append(f"hd{hd}{'+relpos' if relpos else ''}: {s.elapsed_time(e) / 10:.4f} ms")
print(os.environ.get("WM_LIB_NAME", "libwm_b200.so"), "flash_version", os.environ.get("WM_FLASH_VERSION", "default"), " | ".join(res))
