#!/usr/bin/env python
"""Per-class NMS over the 10 k-box dense-herd problem (bench.py measure_nms_10k) on its own."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ctx = bench.Ctx()
print(json.dumps(bench.measure_nms_10k(ctx, cpu_too=True)))
