#!/usr/bin/env python
"""Benchmark of the tile-detection hot path (BASELINE.json metric: 1024^2 tiles/sec through
fft + encoder + decoder + post-process + NMS).

    python bench.py --gpus N --steps K --warmup W                 # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on the host CPU

A step = one batch of synthetic 1024x1024 tiles per GPU through the whole path.  Workload = BASELINE.json
configs[1]: ViT-B detector, bf16 tensor-core math, batch 32 per GPU (weak scaling: every rank runs its own batch,
detections are all-gathered over NCCL).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "wildlifemapper_b200"))

MODEL_CONFIGS = {"vit_b": (768, 12, 12, (2, 5, 8, 11)), "vit_l": (1024, 24, 16, (5, 11, 17, 23)),
                 "vit_h": (1280, 32, 16, (7, 15, 23, 31)), "vit_t": (128, 2, 2, (1,))}
# algorithmic GFLOP per tile (SURVEY.md section 8d / BASELINE.md section 3)
GFLOP_PER_TILE = {"vit_b": 1085.0, "vit_l": 2988.7, "vit_h": 5797.8}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (one `ncu --set full` capture, see profiles/)
NCU_TRAFFIC = {"gemm": 752.8e6}
METRIC = "tiles_per_sec"
UNIT = "tiles/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="vit_b", choices=list(MODEL_CONFIGS))
    ap.add_argument("--batch", type=int, default=32, help="tiles per GPU per step")
    ap.add_argument("--queries", type=int, default=51)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", default="", help="write the per-kernel-family timing JSON here")
    ap.add_argument("--no-graph", action="store_true", help="time the eager (one launch per kernel) path instead of CUDA-graph replay")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": d.get("bf16_tflops_sustained", 1361.2), "gbs": d.get("hbm_gbs", 6550.7), "src": "measured"}
    return {"tflops": 1400.0, "gbs": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_tiles_per_sec(model_type: str, steps: int, warmup: int, queries: int):
    """The reference's algorithm on the host cores: the oracle port (oracle/model.py + oracle/post.py), which
    restates the reference forward with the same ATen CPU kernels the reference itself calls (the reference is
    pure PyTorch and does not travel to the GPU box).  One step = ONE tile (bounded sample of the workload)."""
    import numpy as np
    import torch
    from oracle import model as om
    from oracle import post as opost
    from oracle.weights import make_state_dict, make_tiles
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(model_type, seed=0, num_queries=queries)
    tiles = make_tiles(1, seed=2)
    sizes = np.array([[1024, 1024]])

    def step():
        out = om.forward(sd, model_type, tiles)
        res = opost.postprocess(out["pred_logits"].numpy(), out["pred_boxes"].numpy(), sizes, 0.05)
        for r in res:
            c = r["scores"] > np.float32(0.5)
            opost.nms(r["boxes"][c], r["scores"][c], 0.4)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps, cores, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    tps, s_per_step, cores, threads = cpu_reference_tiles_per_sec(args.model, steps, warmup, args.queries)
    sample = f"{steps} timed + {warmup} warm-up single-tile forwards ({args.model}, fp32, torch CPU, {threads} threads)"
    line = {"impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} detector, 1 tile per step on the host CPU (bounded sample of batch "
                                   f"{args.batch})", "queries": args.queries},
            "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------- B200 arm
def build_model(model_type: str, queries: int, device):
    import torch
    from functools import partial
    from segment_anything.modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, TwoWayTransformer
    from segment_anything.network import MedSAM
    D, depth, heads, glob = MODEL_CONFIGS[model_type]
    torch.manual_seed(0)
    enc = ImageEncoderViT(depth=depth, embed_dim=D, img_size=1024, mlp_ratio=4,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=heads, patch_size=16,
                          qkv_bias=True, use_rel_pos=True, global_attn_indexes=list(glob), window_size=14, out_chans=256)
    pe = PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16)
    dec = MaskDecoder(num_multimask_outputs=queries - 1,
                      transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                      transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256)
    model = MedSAM(image_encoder=enc, mask_decoder=dec, prompt_encoder=pe).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():  # zero-initialised reference parameters get N(0, 0.02) so those paths are live
        for n, p in model.named_parameters():
            if "rel_pos" in n or n.endswith("pos_embed") or n.endswith("in_proj_bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return model.to(device)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from segment_anything.utils.misc import NestedTensor
    from wildlifemapper_b200 import postprocess as pp
    from wildlifemapper_b200 import profiler
    from wildlifemapper_b200.dist import gather_detections

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, Q = args.batch, args.queries
    model = build_model(args.model, Q, dev)
    gen = torch.Generator().manual_seed(2 + rank)
    host_tiles = torch.randn(B, 3, 1024, 1024, generator=gen).pin_memory()
    dev_tiles = host_tiles.to(dev)
    sizes = torch.tensor([[1024, 1024]] * B, device=dev)
    host_out = torch.empty(B, Q, 6).pin_memory()
    host_cnt = torch.empty(B, dtype=torch.int32).pin_memory()
    host_keep = torch.empty(B, Q, dtype=torch.int32).pin_memory()
    host_kcnt = torch.empty(B, dtype=torch.int32).pin_memory()
    enc_eng, dec_eng = model.image_encoder.engine(), model.mask_decoder.transformer.engine()

    def step(tiles):
        with torch.no_grad():
            out = model(NestedTensor(tiles, None), None)
            packed, labels, query, counts = pp.postprocess_packed(out["pred_logits"], out["pred_boxes"], sizes, 0.05)
            keep_idx, keep_cnt = pp.nms_packed(packed, counts, score_thr=0.5, iou_threshold=0.4)
            return gather_detections(packed, counts, keep_idx, keep_cnt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, fn):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n_steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- device-resident throughput (`value`): inputs already in HBM.  Working set per step (>= 2 GB of
    # activations at batch 32) is far larger than the 126 MB L2, so no explicit flush is needed between steps.
    # Two timed passes over the same K steps:
    #   (1) eager, every wm_b200 launch bracketed by CUDA events -> per-kernel-family breakdown and the roofline object;
    #   (2) the product path for a fixed batch shape: the same kernels replayed from a CUDA graph (wildlifemapper_b200/
    #       graph.py) -> `value`.  `--no-graph` reports pass (1) as `value`.
    for _ in range(args.warmup):
        step(dev_tiles)
    l0 = enc_eng.launches + dec_eng.launches
    timer = profiler.start()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_eager = timed(args.steps, lambda: step(dev_tiles))
    profiler.stop()
    fam = timer.summary()
    launches = (enc_eng.launches + dec_eng.launches - l0) + 2 * args.steps  # + postprocess + batched NMS
    graphed = None
    if not args.no_graph:
        from wildlifemapper_b200.graph import GraphedDetector
        graphed = GraphedDetector(model, B, warmup=1)
        graphed.static_in.copy_(dev_tiles)

        def graph_step():
            packed, counts, keep_idx, keep_cnt = graphed.replay()
            return gather_detections(packed, counts, keep_idx, keep_cnt)

        for _ in range(args.warmup):
            graph_step()
        ms = timed(args.steps, graph_step)
    else:
        ms = ms_eager
    clocks = sampler.stop()
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end (`e2e`): pinned host tiles -> H2D -> path -> D2H of the packed detections, every step.  The tile
    # feed is double buffered: the H2D copy of step i+1 runs on a copy stream while step i computes (what a serving
    # loop does); every step still pays one full H2D of its own inputs and one D2H of its own detections inside the
    # timed region, and the caller synchronises on each step's detections.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_buf = [torch.empty_like(dev_tiles), torch.empty_like(dev_tiles)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])  # the step that last read this buffer has finished
            dev_buf[slot].copy_(host_tiles, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step():
        slot = state["i"] & 1
        state["i"] += 1
        prefetch(slot ^ 1)  # next step's tiles (the very first call copies its own tiles below)
        torch.cuda.current_stream().wait_event(ready[slot])
        if graphed is not None:
            graphed.static_in.copy_(dev_buf[slot], non_blocking=True)  # device copy into the captured input (0.1 ms)
            consumed[slot].record()
            packed, counts, keep_idx, keep_cnt = graph_step()
        else:
            packed, counts, keep_idx, keep_cnt = step(dev_buf[slot])
            consumed[slot].record()
        nloc = B  # every rank reads back its own shard's detections
        host_out.copy_(packed[rank * nloc:(rank + 1) * nloc] if world > 1 else packed, non_blocking=True)
        host_cnt.copy_(counts[rank * nloc:(rank + 1) * nloc] if world > 1 else counts, non_blocking=True)
        host_keep.copy_(keep_idx[rank * nloc:(rank + 1) * nloc] if world > 1 else keep_idx, non_blocking=True)
        host_kcnt.copy_(keep_cnt[rank * nloc:(rank + 1) * nloc] if world > 1 else keep_cnt, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller consumes the detections of this step

    for s_ in (0, 1):
        consumed[s_].record()
    prefetch(0)
    e2e_step()
    ms_e2e = timed(args.steps, e2e_step)
    torch.cuda.synchronize()
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = host_tiles.numel() * 4
    d2h = host_out.numel() * 4 + host_cnt.numel() * 4 + host_keep.numel() * 4 + host_kcnt.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    top = max(fam.items(), key=lambda kv: kv[1]["ms"])
    name, r = top
    if r["flop"] > 0:
        achieved = r["flop"] / (r["ms"] / 1e3) / 1e12
        roof = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops"], "traffic": NCU_TRAFFIC.get(name),
                "traffic_note": "DRAM read+write bytes of the qkv-shaped launch (M=131072, N=2304, K=768; algorithmic "
                                "809 MB) from ncu --set full, profiles/r01z_ncu_gemm_qkv_summary.txt (the tail of the output is still in L2 when the kernel ends)",
                "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside a long step)",
                "share_of_step": r["ms"] / ms_eager, "launches_per_step": r["launches"] / args.steps,
                "avg_launch_ms": r["ms"] / r["launches"],
                "measured_in": "event-instrumented eager pass over the same K steps (one CUDA-event pair per launch)"}
    else:
        achieved = r["byte"] / (r["ms"] / 1e3) / 1e9
        roof = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peaks["gbs"], "unit": "GB/s",
                "frac": achieved / peaks["gbs"], "traffic": None, "peak_source": peaks["src"],
                "share_of_step": r["ms"] / ms_eager}
    breakdown = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                     "tflops": (v["flop"] / (v["ms"] / 1e3) / 1e12) if v["flop"] else None,
                     "gbs": (v["byte"] / (v["ms"] / 1e3) / 1e9) if v["byte"] else None}
                 for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    if args.breakdown:
        with open(args.breakdown, "w") as f:
            json.dump({"ms_per_step": ms_eager / args.steps, "batch": B, "model": args.model, "families": breakdown}, f, indent=1)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        tps, spt, cores, threads = cpu_reference_tiles_per_sec(args.model, 2, 1, Q)
        cpu = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 timed + 1 warm-up single-tile forwards of the oracle port ({args.model}, fp32, {threads} threads)"}
    gf = GFLOP_PER_TILE.get(args.model)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} detector (fft + encoder + decoder + PostProcess + NMS), batch {B} "
                                   f"synthetic 1024x1024 tiles per GPU, {Q} queries", "model_type": args.model,
                       "batch_per_gpu": B, "parallelism": f"tile-sharded x{world}, NCCL all-gather of detections",
                       "l2": "inputs and activations per step exceed L2 (no flush needed)",
                       "launch": "eager (one launch per kernel)" if graphed is None else "CUDA-graph replay of the same kernels"},
            "eager": {"value": world * B * args.steps / (ms_eager / 1e3), "unit": UNIT, "ms_per_step": ms_eager / args.steps},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "model_tflops": (value * gf / 1e3) if gf else None,
            "model_frac_of_bf16_peak": (value / world * gf / 1e3 / peaks["tflops"]) if gf else None,
            "breakdown_ms_per_step": {k: round(v["ms_per_step"], 3) for k, v in breakdown.items()}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """Write the ONE JSON line to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # Libraries print to fd 1 behind Python's back (NCCL's version banner under torchrun): keep stdout for the JSON line only.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
