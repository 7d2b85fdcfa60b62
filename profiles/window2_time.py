#!/usr/bin/env python
"""Launch time of the windowed-attention kernel (ViT-B shapes, batch 32) vs the logit scale of the synthetic input."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
B, H, hd = (64, 16, 80) if len(sys.argv) > 1 and sys.argv[1] == "vit_h" else (32, 12, 64)  # ViT-H: head dim 80 (window3 kernel)
D = H * hd
out = torch.empty(B * 4096, D, device="cuda", dtype=torch.bfloat16)
for sc in (0.05, 0.3, 1.0):
    qkv = (torch.randn(B * 4096, 3 * D, device="cuda") * sc).to(torch.bfloat16)
    table = (torch.randn(64, hd, device="cuda") * 0.3 * sc).to(torch.bfloat16)
    for _ in range(3):
        ops.attn_window(qkv, table, out, H, 1 / math.sqrt(hd))
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.attn_window(qkv, table, out, H, 1 / math.sqrt(hd))
    e.record(); torch.cuda.synchronize()
    print(f"scale {sc}: {s.elapsed_time(e) / 10:.4f} ms per launch")
