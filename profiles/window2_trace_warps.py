#!/usr/bin/env python
"""Per-warp skew of tile 0 in window2 (diagnostics build with -DWM_F3_TRACE -DWM_W2_TRACE_WARPS)."""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
from wildlifemapper_b200 import lib
B, H, hd = 32, 12, 64
D = H * hd
SC = float(os.environ.get("W2_SCALE", "1.0"))  # 1.0: harsh logits (frequent reference-maximum raises); 0.3: model-like
qkv = (torch.randn(B * 4096, 3 * D, device="cuda") * SC).to(torch.bfloat16)
out = torch.empty(B * 4096, D, device="cuda", dtype=torch.bfloat16)
table = (torch.randn(64, hd, device="cuda") * 0.3 * SC).to(torch.bfloat16)
for _ in range(3):
    ops.attn_window(qkv, table, out, H, 1 / math.sqrt(hd))
torch.cuda.synchronize()
buf = np.zeros((3, 64, 8), dtype=np.uint64)
lib.call("wm_debug_window_trace", buf.ctypes.data)
g = buf[2].astype(np.int64)
for n in range(20, 28):
    base = g[n, :4].min()
    print(f"n={n} pass start (warp 0..3) {(g[n, :4] - base).tolist()}  P arrive {(g[n, 4:8] - base).tolist()}  pass len {(g[n, 4:8] - g[n, :4]).tolist()}")
sl = slice(8, 56)
print("mean pass length per warp", (g[sl, 4:8] - g[sl, :4]).mean(0).round().tolist())
print("mean arrive skew (last - first)", float((g[sl, 4:8].max(1) - g[sl, 4:8].min(1)).mean()))
