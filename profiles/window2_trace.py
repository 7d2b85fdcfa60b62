#!/usr/bin/env python
"""Timeline of one window-attention (v2) CTA (diagnostics build with -DWM_F3_TRACE): python profiles/window2_trace.py"""
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
for _ in range(2):
    ops.attn_window(qkv, table, out, H, 1 / math.sqrt(hd))
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); ops.attn_window(qkv, table, out, H, 1 / math.sqrt(hd)); e.record(); torch.cuda.synchronize()
print("launch ms", s.elapsed_time(e), "items per CTA", B * 25 * H / 148)
buf = np.zeros((3, 64, 8), dtype=np.uint64)
lib.call("wm_debug_window_trace", buf.ctypes.data)
rel = buf.astype(np.int64) - int(buf[buf > 0].min())
print("softmax: 0 reached S wait, 1 S ready, 2 bias in regs, 3 P arrived, 4 O seen, 5 s_free arrived, 6 store issued")
print("mma (per tile t: 4t+): 0 P seen, 1 PV issued, 2 S' deps ready, 3 S' issued")
for n in range(20, 24):
    print(f"n={n} sm0 {rel[0, n, :7].tolist()}\n      sm1 {rel[1, n, :7].tolist()}\n      mma {rel[2, n, :8].tolist()}")
sl = slice(8, 56)
m = lambda a: float(np.mean(a))
print("item period:", m(np.diff(rel[0, 8:57, 0])))
for t in (0, 1):
    r = rel[t]
    print(f"tile {t}: S wait {m(r[sl,1]-r[sl,0]):.0f} | tables {m(r[sl,2]-r[sl,1]):.0f} | pass {m(r[sl,3]-r[sl,2]):.0f} | O wait {m(r[sl,4]-r[sl,3]):.0f} "
          f"| O read {m(r[sl,5]-r[sl,4]):.0f} | scale+store {m(r[sl,6]-r[sl,5]):.0f}")
g = rel[2]
for t in (0, 1):
    o = 4 * t
    print(f"mma t{t}: PV issue {m(g[sl,o+1]-g[sl,o]):.0f} | wait for S' deps {m(g[sl,o+2]-g[sl,o+1]):.0f} | S' issue {m(g[sl,o+3]-g[sl,o+2]):.0f}")
print(f"mma: S'0 issued -> P1 seen {m(g[sl,4]-g[sl,3]):.0f}; S'1 issued -> next P0 seen {m(g[9:57,0]-g[8:56,7]):.0f}")
print(f"P arrive (t0) -> seen by mma {m(g[9:57,0]-rel[0][8:56,3]):.0f}; t1 {m(g[9:57,4]-rel[1][8:56,3]):.0f}")
print(f"PV issued -> O seen: t0 {m(rel[0][8:56,4]-g[9:57,1]):.0f}; t1 {m(rel[1][8:56,4]-g[9:57,5]):.0f}")
print(f"S' issued -> S ready: t0 {m(rel[0][sl,1]-g[sl,3]):.0f}; t1 {m(rel[1][sl,1]-g[sl,7]):.0f}")
