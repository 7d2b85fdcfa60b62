#!/usr/bin/env python
"""Measure the survey front-end / merge / COCO packing kernels on the B200.
  tiles_from_u8: algorithmic bytes = T x 1024^2 x (3 read + 12 written), against the measured HBM copy bandwidth
  merge + per-class NMS + COCO packing: latency for one survey image (24 tiles x Q rows)
  SurveyDetector end to end: uint8 3648 x 5472 image in pinned host memory -> COCO records on the host
Usage (GPU box): python profiles/frontend_bench.py [vit_b] > gpurun_out/frontend_bench.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "wildlifemapper_b200"))
from wildlifemapper_b200 import survey  # noqa: E402
from bench import build_model, load_peaks  # noqa: E402

dev = torch.device("cuda")
model_type = sys.argv[1] if len(sys.argv) > 1 else "vit_b"
H, W = 3648, 5472


def timeit(fn, iters=10, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()  # evict L2 (the flush buffer is larger than L2)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


rng = np.random.default_rng(0)
img_host = torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).pin_memory()
img = img_host.to(dev)
origins_host = survey.plan_tiles(H, W, 1024, 128)
origins = torch.tensor(origins_host, dtype=torch.int32, device=dev)
T = origins.shape[0]
out = torch.empty(T, 3, 1024, 1024, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = load_peaks()
res = {"image": [H, W], "tiles": T}
ms = timeit(lambda: survey.tiles_from_u8(img, origins, out=out), flush=flush)
byts = T * 1024 * 1024 * 15
res["tiles_from_u8"] = {"ms": ms, "algorithmic_bytes": byts, "gbs": byts / ms / 1e6, "peak_gbs": peaks["gbs"],
                        "frac": byts / ms / 1e6 / peaks["gbs"], "note": "L2 flushed between iterations"}

for Q in (51, 900):
    from oracle.frontend import make_tile_detections  # seeded synthetic PostProcess rows (input generator only)
    packed, counts = make_tile_detections(T, Q, seed=3)
    counts[:] = Q
    packed_d, counts_d = torch.from_numpy(packed).to(dev), torch.from_numpy(counts).to(dev)

    def merge_all():
        m = survey.merge_tile_detections(packed_d, counts_d, origins, 0.5, 0.4, True)
        return survey.coco_records(m["boxes"], m["scores"], m["labels"], m["keep"])

    ms = timeit(merge_all)
    m = survey.merge_tile_detections(packed_d, counts_d, origins, 0.5, 0.4, True)
    res[f"merge_nms_pack_q{Q}"] = {"ms": ms, "candidates": int(m["scores"].shape[0]), "kept": int(m["keep"].shape[0]),
                                   "note": "merge + per-class NMS (2 host syncs for the data-dependent lengths) + COCO packing"}

model = build_model(model_type, 51, dev)
det = survey.SurveyDetector(model, batch=T)


def e2e():
    m = det(img_host)
    xywh, cat = survey.coco_records(m["boxes"], m["scores"], m["labels"], m["keep"])
    return survey.coco_dicts(1, xywh, cat)


ms = timeit(e2e, iters=5)
res["survey_e2e"] = {"model": model_type, "ms_per_image": ms, "tiles_per_sec": T / ms * 1e3, "images_per_sec": 1e3 / ms,
                     "h2d_bytes_per_image": H * W * 3,
                     "note": "pinned uint8 image -> H2D -> tiles_from_u8 -> model (eager) -> PostProcess -> merge + per-class NMS -> COCO dicts on the host"}
print(json.dumps(res))
