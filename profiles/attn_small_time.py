#!/usr/bin/env python
"""Launch times of the decoder attention shapes (batch 32, 8 heads): self (51x51, hd 32), tokens->image (51x4096, hd 16),
image->tokens (4096x51, hd 16)."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
B, H = 32, 8
for name, Tq, Tk, hd in (("self 51x51 hd32", 51, 51, 32), ("t2i 51x4096 hd16", 51, 4096, 16), ("i2t 4096x51 hd16", 4096, 51, 16)):
    D = H * hd
    q = torch.randn(B * Tq, D, device="cuda").to(torch.bfloat16)
    k = torch.randn(B * Tk, D, device="cuda").to(torch.bfloat16)
    v = torch.randn(B * Tk, D, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * Tq, D, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.attn_small(q, k, v, o, B, H, Tq, Tk, hd, 1 / math.sqrt(hd))
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        f()
    e.record(); torch.cuda.synchronize()
    print(f"{os.environ.get('WM_LIB_NAME', 'libwm_b200.so')} {name}: {s.elapsed_time(e) / 20 * 1e3:.1f} us")
