#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i): headline metrics per captured launch + top stall reasons / hottest SASS lines.
Usage: python profiles/ncu_summary.py gpurun_out/prof_X.ncu-rep [n_lines]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.avg.per_cycle_active"]
idx = [(i, h) for i, h in enumerate(hdr) if h in want]
for r in rows[2:]:
    print({h: r[i][:48] for i, h in idx})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    h = rows[hi[0]]
    body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
    ci = {n: i for i, n in enumerate(h)}
    f = lambda r, n: float(r[ci[n]]) if r[ci[n]] not in ("", "-") else 0.0
    tot = sum(f(r, "# Samples") for r in body) or 1.0
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = sorted(((sum(f(r, n) for r in body), n) for n in stalls), reverse=True)
    print("total samples", tot, "sass instructions", len(body))
    print("stalls:", ", ".join(f"{n[6:]} {100 * v / tot:.1f}%" for v, n in agg[:9]))
    for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:nl]:
        top = max(stalls, key=lambda n: f(r, n))
        print(f"{100 * f(r, '# Samples') / tot:5.1f}% ex={f(r, 'Instructions Executed'):11.0f} {top[6:]:18s} {r[ci['Source']][:100]}")
