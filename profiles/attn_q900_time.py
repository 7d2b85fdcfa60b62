#!/usr/bin/env python
"""Decoder attention at 900 queries (dense herd), batch 32: the warp-level MMA kernel on the native 16 / 32-wide heads vs the
tcgen05 flash kernel on head-padded (64-wide) rows."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
B, H = 32, 8


def timeit(f, n=10):
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


for name, Tq, Tk, hd in (("self 900x900 hd32", 900, 900, 32), ("t2i 900x4096 hd16", 900, 4096, 16), ("i2t 4096x900 hd16", 4096, 900, 16)):
    D = H * hd
    q = torch.randn(B * Tq, D, device="cuda").to(torch.bfloat16)
    k = torch.randn(B * Tk, D, device="cuda").to(torch.bfloat16)
    v = torch.randn(B * Tk, D, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * Tq, D, device="cuda", dtype=torch.bfloat16)
    t_mma = timeit(lambda: ops.attn_small(q, k, v, o, B, H, Tq, Tk, hd, 1 / math.sqrt(hd)))
    qp = torch.randn(B * Tq, 512, device="cuda").to(torch.bfloat16)
    kp = torch.randn(B * Tk, 512, device="cuda").to(torch.bfloat16)
    vp = torch.randn(B * Tk, 512, device="cuda").to(torch.bfloat16)
    op = torch.empty(B * Tq, 512, device="cuda", dtype=torch.bfloat16)
    t_fl = timeit(lambda: ops.attn_flash(qp, 0, kp, 0, vp, 0, None, op, B, 8, Tq, Tk, 64, 1 / math.sqrt(hd)))
    print(f"{name}: mma {t_mma * 1e3:.1f} us | flash (head-padded) {t_fl * 1e3:.1f} us")
