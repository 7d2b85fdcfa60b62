#!/usr/bin/env python
"""Timeline of one flash-v4 CTA (diagnostics build with -DWM_F3_TRACE -rdc): python profiles/flash4_trace.py [hd] [relpos]"""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
from wildlifemapper_b200 import lib
hd = int(sys.argv[1]) if len(sys.argv) > 1 else 64
relpos = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B, H, T = 8, 12 if hd == 64 else 8, 4096
D = H * hd
qkv = torch.randn(B * T, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
table = (torch.randn(256, hd, device="cuda") * 0.3).to(torch.bfloat16) if relpos else None
for _ in range(2):
    ops.attn_flash(qkv, 0, qkv, D, qkv, 2 * D, table, out, B, H, T, T, hd, 1 / math.sqrt(hd))
torch.cuda.synchronize()
buf = np.zeros((3, 64, 8), dtype=np.uint64)
lib.call("wm_debug_flash_trace", buf.ctypes.data)
rel = buf.astype(np.int64) - int(buf[buf > 0].min())
print("softmax: 0 reached wait, 1 S ready, 2 pass done, 3 arrived | mma: t0: 0 reached wait 1 P seen 2 PV issued 3 S issued; t1: 4..7")
for j in range(40, 46):
    print(f"j={j}  sm0 {rel[0, j, :4].tolist()}  sm1 {rel[1, j, :4].tolist()}  mma {rel[2, j, :8].tolist()}")
sl = slice(16, 60)
m = lambda a: float(np.mean(a))
print("step period:", m(np.diff(rel[0, 16:61, 0])))
for t in (0, 1):
    r = rel[t]
    print(f"tile {t}: S wait {m(r[sl,1]-r[sl,0]):.0f}  pass {m(r[sl,2]-r[sl,1]):.0f}  st-wait+arrive {m(r[sl,3]-r[sl,2]):.0f}")
g = rel[2]
print(f"mma t0: P wait {m(g[sl,1]-g[sl,0]):.0f} PV issue {m(g[sl,2]-g[sl,1]):.0f} S issue {m(g[sl,3]-g[sl,2]):.0f} | t1: P wait {m(g[sl,5]-g[sl,4]):.0f} PV issue {m(g[sl,6]-g[sl,5]):.0f} S issue {m(g[sl,7]-g[sl,6]):.0f}")
print(f"P arrive(t0,j) -> seen {m(g[sl,1]-rel[0][sl,3]):.0f}; t1 {m(g[sl,5]-rel[1][sl,3]):.0f}")
