#!/usr/bin/env python
"""Per-step cost and per-CTA fixed cost of the head-dim-64 flash kernels: launch time vs key count (no rel-pos: any Tk) and the
rel-pos launch at 4096 keys, for every selectable generation, interleaved inside ONE process (same box, same clocks).
    python profiles/flash_time2.py [batch] [versions, e.g. 4,7]"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
from wildlifemapper_b200 import lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
versions = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4", "7"])]
H, hd, T = 12, 64, 4096
D = H * hd
qkv = (torch.randn(B * T, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
table = (torch.randn(256, hd, device="cuda") * 0.05).to(torch.bfloat16)


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


for rep in range(2):
    for v in versions:
        lib.call("wm_set_flash_version", v)
        row = []
        for Tk in (1024, 2048, 4096):
            k = qkv[: B * Tk]  # image b's keys start at b * Tk: any rows will do for timing
            ms = timeit(lambda: ops.attn_flash(qkv, 0, k, D, k, 2 * D, None, out, B, H, T, Tk, hd, 1 / math.sqrt(hd)))
            row.append((Tk, ms))
        rrow = []
        for Tk in (2048, 4096):
            k = qkv[: B * Tk]
            rrow.append(timeit(lambda: ops.attn_flash(qkv, 0, k, D, k, 2 * D, table, out, B, H, T, Tk, hd, 1 / math.sqrt(hd))))
        rel = rrow[1]
        rel_step = (rrow[1] - rrow[0]) / 32 / (((T // 256) * H * B) / 148) * 1e6
        rel_fixed = (rrow[0] - (rrow[1] - rrow[0])) / (((T // 256) * H * B) / 148) * 1e6
        ctas = (T // 256) * H * B
        waves = ctas / 148
        per_step = (row[2][1] - row[1][1]) / 32 / waves * 1e6  # ns per 64-key step of one CTA
        fixed = (row[1][1] - (row[2][1] - row[1][1])) / waves * 1e6  # ns per CTA outside the steps
        print(f"v{v}: " + " ".join(f"Tk{t}={m:.3f}ms" for t, m in row) + f" relpos4096={rel:.3f}ms | per step {per_step:.0f} ns, fixed per CTA {fixed:.0f} ns "
              f"| relpos: per step {rel_step:.0f} ns, fixed per CTA {rel_fixed:.0f} ns")
