#!/usr/bin/env python
"""Is the rel-pos flash launch power bound?  Times it (a) back to back (SM clock sampled with NVML while it runs) and (b) as
single launches separated by idle gaps (clocks recover), for each selectable generation.
    python profiles/flash_power.py [batch] [versions]"""
import math, os, sys, time, threading
import torch
import pynvml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wildlifemapper_b200.ops import ops
from wildlifemapper_b200 import lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
versions = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["7", "9"])]
H, hd, T = 12, 64, 4096
D = H * hd
qkv = (torch.randn(B * T, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
table = (torch.randn(256, hd, device="cuda") * 0.05).to(torch.bfloat16)
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)


def launch():
    ops.attn_flash(qkv, 0, qkv, D, qkv, 2 * D, table, out, B, H, T, T, hd, 1 / math.sqrt(hd))


def sampled(n):
    clocks, power, stop = [], [], [False]

    def poll():
        while not stop[0]:
            clocks.append(pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM))
            power.append(pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000)
            time.sleep(0.01)
    th = threading.Thread(target=poll)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20):
        launch()
    torch.cuda.synchronize()
    th.start()
    s.record()
    for _ in range(n):
        launch()
    e.record()
    torch.cuda.synchronize()
    stop[0] = True
    th.join()
    clocks.sort(); power.sort()
    return s.elapsed_time(e) / n, clocks[len(clocks) // 2], power[len(power) // 2]


def isolated(n, gap):
    ts = []
    for _ in range(n):
        time.sleep(gap)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); launch(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for rep in range(2):
    for v in versions:
        lib.call("wm_set_flash_version", v)
        ms, clk, pw = sampled(400)
        med, best = isolated(20, 0.05)
        print(f"v{v}: back to back {ms:.3f} ms (SM {clk} MHz, {pw:.0f} W) | isolated median {med:.3f} ms, best {best:.3f} ms")
