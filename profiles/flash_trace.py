#!/usr/bin/env python
"""Timeline of one flash-v3 CTA (diagnostics build: WM_LIB_NAME=libwm_b200_dbg.so python profiles/flash_trace.py [hd] [relpos])."""
import ctypes, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
from wildlifemapper_b200 import lib
hd = int(sys.argv[1]) if len(sys.argv) > 1 else 64
relpos = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B, H, T = 8, 12 if hd == 64 else 8, 4096
D = H * hd
qkv = (torch.randn(B * T, 3 * D, device="cuda")).to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
table = (torch.randn(256, 64, device="cuda") * 0.3).to(torch.bfloat16) if relpos else None
for _ in range(2):
    ops.attn_flash(qkv, 0, qkv, D, qkv, 2 * D, table, out, B, H, T, T, hd, 1 / math.sqrt(hd))
torch.cuda.synchronize()
buf = np.zeros((3, 64, 8), dtype=np.uint64)
lib.call("wm_debug_flash_trace", buf.ctypes.data)
t0 = int(buf[buf > 0].min())
rel = buf.astype(np.int64) - t0
print(f"hd={hd} relpos={relpos}; SM cycles relative to the first event")
print("softmax: 0 S_ready 1 first chunk loaded 3 turn 4 P_issued 5 P_stored | mma: 0 P0_seen 1 PV0_issued 2 S0_issued 3 P1_seen 4 PV1_issued 5 S1_issued")
for j in range(26, 31):
    print(f"j={j:2d}  sm0 {rel[0, j, [0,1,3,4,5]].tolist()}  sm1 {rel[1, j, [0,1,3,4,5]].tolist()}  mma {rel[2, j, :6].tolist()}")
sl = slice(8, 30)
nx = slice(9, 31)
m = lambda a: float(np.mean(a))
print("tile-0 period (cycles per key tile):", m(np.diff(rel[0, 8:31, 0])))
for t in (0, 1):
    r = rel[t]
    print(f"tile {t}: load {m(r[sl,1]-r[sl,0]):.0f}  wait-turn {m(r[sl,3]-r[sl,1]):.0f}  softmax pass {m(r[sl,4]-r[sl,3]):.0f}  "
          f"st-wait {m(r[sl,5]-r[sl,4]):.0f}  P_stored->S_ready(next) {m(r[nx,0]-r[sl,5]):.0f}")
g = rel[2]
print(f"mma: P0 stored->seen {m(g[sl,0]-rel[0][sl,5]):.0f}  PV0 issue {m(g[sl,1]-g[sl,0]):.0f}  S0 issue {m(g[sl,2]-g[sl,1]):.0f}  S0 issued->S_ready {m(rel[0][nx,0]-g[sl,2]):.0f} | "
      f"P1 stored->seen {m(g[sl,3]-rel[1][sl,5]):.0f}  PV1 issue {m(g[sl,4]-g[sl,3]):.0f}  S1 issue {m(g[sl,5]-g[sl,4]):.0f}  S1 issued->S_ready {m(rel[1][nx,0]-g[sl,5]):.0f}")
