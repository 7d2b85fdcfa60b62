#!/usr/bin/env python
"""The three transposes of the step (batch 32) against the HBM copy bandwidth."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
for name, B, R, C, dt in (("lowpass p1", 32, 1024, 2048, torch.bfloat16), ("hfc scramble", 32, 1024, 4096, torch.bfloat16), ("NCHW features", 32, 4096, 256, torch.float32)):
    x = torch.randn(B, R, C, device="cuda").to(dt)
    o = torch.empty(B, C, R, device="cuda", dtype=dt)
    for _ in range(3):
        ops.transpose(x, o)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.transpose(x, o)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"{name}: {ms * 1e3:.1f} us, {2 * x.numel() * x.element_size() / ms / 1e6:.0f} GB/s")
