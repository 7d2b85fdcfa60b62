#!/usr/bin/env python
"""Launch time of the 14x14 window attention (head dim 64 and 80) at the bench shapes: python profiles/window_time.py [batch]"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        f()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for H, hd in ((12, 64), (16, 64), (16, 80)):
    D = H * hd
    qkv = (torch.randn(B * 4096, 3 * D, device="cuda") * 0.3).to(torch.bfloat16)
    out = torch.empty(B * 4096, D, device="cuda", dtype=torch.bfloat16)
    table = (torch.randn(64, hd, device="cuda") * 0.1).to(torch.bfloat16)
    ms = [timeit(lambda: ops.attn_window(qkv, table, out, H, 1 / math.sqrt(hd))) for _ in range(3)]
    flops = B * 25 * H * 4 * 196 * 196 * hd
    print(f"window H={H} hd={hd} B={B}: " + " ".join(f"{m:.4f}" for m in ms) + f" ms  ({flops / min(ms) / 1e9:.0f} TFLOP/s)")
    del qkv, out
