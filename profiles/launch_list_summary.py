#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py (profiles/run_launch_list.sh):
per-family share of the device time of ONE pass of the step (cold-cache, serialised launches: compare the SHARES with
bench.py's event breakdown, not the absolute times).  Usage: python profiles/launch_list_summary.py in.csv out_prefix"""
import csv, json, sys
src, prefix = sys.argv[1], sys.argv[2]
rows = []
hdr = None
for r in csv.reader(open(src)):
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((d["Kernel Name"], float(d["Metric Value"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(d.get("Metric Unit", "ns"), 1e-6)))
fam_of = lambda n: next((f for f in ("gemm", "flash", "window", "layernorm", "attn_small", "transpose", "patchify", "hfc_finalize",
                                     "add_cast", "postprocess", "nms", "sigmoid", "rank", "topk") if f in n), "other")
# one pass = from a patchify launch (first kernel of the step) to the next one
starts = [i for i, (n, _) in enumerate(rows) if "patchify" in n]
one = rows[starts[1]:starts[2]] if len(starts) > 2 else rows
agg = {}
for n, ms in one:
    f = fam_of(n)
    agg.setdefault(f, [0, 0.0])
    agg[f][0] += 1
    agg[f][1] += ms
tot = sum(v[1] for v in agg.values())
out = {"launches": len(one), "sum_ms": round(tot, 3),
       "share": {k: round(v[1] / tot, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
       "ms": {k: round(v[1], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
       "count": {k: v[0] for k, v in agg.items()}}
json.dump(out, open(prefix + "_summary.json", "w"), indent=0)
with open(prefix + ".csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "gpu__time_duration_ms"])
    for n, ms in one:
        w.writerow([n[:120], f"{ms:.6f}"])
print(json.dumps(out))
