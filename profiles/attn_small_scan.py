import math, os, sys, torch
sys.path.insert(0, "/root/repo")
from wildlifemapper_b200.ops import ops
B, H = 32, 8
for name, Tq, Tk, hd in (("t2i", 51, 1024, 16), ("t2i", 51, 2048, 16), ("t2i", 51, 4096, 16), ("t2i", 51, 8192, 16), ("t2i-nocluster", 300, 4096, 16), ("i2t", 4096, 51, 16), ("i2t", 4096, 64, 16), ("i2t", 2048, 51, 16)):
    D = H * hd
    q = torch.randn(B * Tq, D, device="cuda").to(torch.bfloat16)
    k = torch.randn(B * Tk, D, device="cuda").to(torch.bfloat16)
    v = torch.randn(B * Tk, D, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * Tq, D, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.attn_small(q, k, v, o, B, H, Tq, Tk, hd, 1 / math.sqrt(hd))
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): f()
    e.record(); torch.cuda.synchronize()
    print(f"{name} {Tq}x{Tk} hd{hd}: {s.elapsed_time(e) / 20 * 1e3:.1f} us")
