#!/usr/bin/env python
"""Benchmark of the tile-detection hot path (BASELINE.json metric: 1024^2 tiles/sec through
fft + encoder + decoder + post-process + NMS).

    python bench.py --gpus N --steps K --warmup W                 # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference itself on the host CPU

A step = one batch of synthetic 1024x1024 tiles per GPU through the whole path.  Headline workload = BASELINE.json
configs[1]: ViT-B detector, bf16 tensor-core math, batch 32 per GPU (weak scaling: every rank runs its own batch,
detections are all-gathered over NCCL).  After the headline the same invocation runs short passes of the other
BASELINE.json configurations (ViT-L batch 32, ViT-H batch 64, dense herd: 900 queries + per-class NMS over 10 k boxes) and
reports them under "configs" / "dense_herd" of the same JSON line.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_CONFIGS = {"vit_b": (768, 12, 12, (2, 5, 8, 11)), "vit_l": (1024, 24, 16, (5, 11, 17, 23)),
                 "vit_h": (1280, 32, 16, (7, 15, 23, 31)), "vit_t": (128, 2, 2, (1,))}
# algorithmic GFLOP per tile (SURVEY.md section 8d / BASELINE.md section 3); 900 queries add 16 GF of decoder work
GFLOP_PER_TILE = {"vit_b": 1085.0, "vit_l": 2988.7, "vit_h": 5797.8}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel family's largest launch (one
# `ncu --set full` capture of this bench command, see profiles/): not measurable inside the run itself
NCU_TRAFFIC = {"gemm": 752.8e6}
METRIC = "tiles_per_sec"
UNIT = "tiles/s"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


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
    ap.add_argument("--no-extras", action="store_true", help="headline only (skip the ViT-L / ViT-H / dense-herd passes)")
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


def dense_herd_boxes(n: int = 10000, seed: int = 3):
    """SURVEY.md section 8(d) NMS stress generator: centres U(0,1024)^2, w,h = exp(N(ln 32, 0.4)) px clipped to the image,
    labels U{0..6}, DISTINCT fp32 scores (randperm + 0.5) / n."""
    import numpy as np
    rng = np.random.default_rng(seed)
    cx, cy = rng.uniform(0, 1024, n), rng.uniform(0, 1024, n)
    w, h = np.exp(rng.normal(np.log(32.0), 0.4, n)), np.exp(rng.normal(np.log(32.0), 0.4, n))
    boxes = np.stack([np.clip(cx - w / 2, 0, 1024), np.clip(cy - h / 2, 0, 1024), np.clip(cx + w / 2, 0, 1024),
                      np.clip(cy + h / 2, 0, 1024)], -1).astype(np.float32)
    scores = ((rng.permutation(n) + 0.5) / n).astype(np.float32)
    labels = rng.integers(0, 7, n).astype(np.int64)
    return boxes, scores, labels


# ----------------------------------------------------------------------------- reference arm (CPU)
def _reference_step_fn(model_type: str, queries: int):
    """-> (step(), kind, note).  kind "reference": the UNMODIFIED reference package (baseline/_ref/segment_anything, copied
    from /root/reference by __graft_entry__.build(); pure Python, no native code) -- MedSAM.forward + PostProcess +
    torchvision.ops.nms exactly as inference.py / visualize_prediction.py call them.  kind "port": the oracle restatement
    (oracle/model.py + oracle/post.py, same ATen CPU kernels) when the copy is absent or the query count is not the
    reference's hard-coded 51."""
    import numpy as np
    import torch
    from oracle.weights import make_state_dict, make_tiles
    tiles = make_tiles(1, seed=2)
    sd = make_state_dict(model_type, seed=0, num_queries=queries)
    if queries == 51 and os.path.isdir(os.path.join(REF_DIR, "segment_anything")):
        try:
            import types
            for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
                del sys.modules[k]
            sys.path.insert(0, REF_DIR)
            try:
                import segment_anything as rsa
                from segment_anything.network import MedSAM as RefMedSAM
                from segment_anything.utils.misc import NestedTensor as RefNested
                import torchvision
            finally:
                sys.path.remove(REF_DIR)
            assert os.path.realpath(rsa.__file__).startswith(os.path.realpath(REF_DIR)), rsa.__file__
            ns = types.SimpleNamespace(set_cost_class=1, set_cost_bbox=5, set_cost_giou=2, bbox_loss_coef=5,
                                       giou_loss_coef=2, eos_coef=0.1, device="cpu")
            sam, _crit, post = rsa.sam_model_registry[model_type](checkpoint=None, args=ns)
            model = RefMedSAM(sam.image_encoder, sam.mask_decoder, sam.prompt_encoder).eval()
            model.load_state_dict(sd, strict=True)
            sizes = torch.tensor([[1024, 1024]])
            box = np.array([[0, 0, 1024, 1024]])

            def step():
                with torch.no_grad():
                    out = model(RefNested(tiles, None), box)
                    res = post["bbox"](out, sizes)
                    for r in res:  # visualize_prediction.py:150-154
                        if r["scores"].numel():
                            c = r["scores"] > 0.5
                            torchvision.ops.nms(r["boxes"][c], r["scores"][c], 0.4)
                return out

            return step, "reference", "unmodified reference package from baseline/_ref (MedSAM.forward + PostProcess + torchvision nms)"
        except Exception as e:  # fall through to the port, but say why
            print(f"[bench] reference package unusable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    from oracle import model as om
    from oracle import post as opost
    sizes_np = np.array([[1024, 1024]])

    def step():
        out = om.forward(sd, model_type, tiles)
        res = opost.postprocess(out["pred_logits"].numpy(), out["pred_boxes"].numpy(), sizes_np, 0.05)
        for r in res:
            c = r["scores"] > np.float32(0.5)
            opost.nms(r["boxes"][c], r["scores"][c], 0.4)
        return out

    return step, "port", "oracle restatement of the reference forward (same ATen CPU kernels)"


def cpu_reference_tiles_per_sec(model_type: str, steps: int, warmup: int, queries: int):
    """The reference on the host cores, ONE tile per step (bounded sample of the workload)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, note = _reference_step_fn(model_type, queries)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"tps": steps / dt, "s_per_step": dt / steps, "cores": cores, "threads": torch.get_num_threads(), "kind": kind,
            "note": note}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    r = cpu_reference_tiles_per_sec(args.model, steps, warmup, args.queries)
    sample = (f"{steps} timed + {warmup} warm-up single-tile forwards ({args.model}, fp32, torch CPU, {r['threads']} threads): "
              f"{r['note']}")
    line = {"impl": "reference", "metric": METRIC, "value": r["tps"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} detector, 1 tile per step on the host CPU (bounded sample of batch "
                                   f"{args.batch})", "queries": args.queries},
            "cpu_baseline": {"value": r["tps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": r["tps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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


class Ctx:
    """Process-wide state of the B200 arm (device, ranks, timing helpers)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        # torchrun exports OMP_NUM_THREADS=1: the random initialisation of the models on the host would take minutes
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, self.world)))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, n_steps, fn):
        """EXACTLY n_steps calls bracketed by barrier + synchronize; CUDA events on the launching stream; max over ranks."""
        torch = self.torch
        self.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n_steps):
            fn()
        e.record()
        self.barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item()


def measure_config(ctx: Ctx, model_type: str, B: int, Q: int, steps: int, warmup: int, *, per_class: bool = False,
                   use_graph: bool = True, want_e2e: bool = True, sample_clocks: bool = False):
    """One workload through the whole path on every rank: (1) eager pass with every wm_b200 launch bracketed by CUDA
    events -> per-kernel-family breakdown + roofline object; (2) the product path for a fixed batch shape: the same
    kernels (and the NCCL all-gather of the detections) replayed from one CUDA graph -> `value`; (3) end to end from
    pinned host tiles with the detections read back every step -> `e2e`."""
    torch = ctx.torch
    from segment_anything.utils.misc import NestedTensor
    from wildlifemapper_b200 import postprocess as pp
    from wildlifemapper_b200 import profiler
    from wildlifemapper_b200.dist import gather_buffer
    from wildlifemapper_b200.graph import GraphedDetector
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    # NMS candidates: the reference's visualisation keeps score > 0.5 before NMS (visualize_prediction.py:150); random-init
    # heads score ~0.14, so that threshold would leave the NMS kernel with nothing to do.  The bench lowers the NMS score
    # threshold to the PostProcess confidence threshold (0.05): every PostProcess row is an NMS candidate.
    conf_thr, nms_thr, iou_thr = 0.05, 0.05, 0.4
    model = build_model(model_type, Q, dev)
    gen = torch.Generator().manual_seed(2 + rank)
    host_tiles = torch.randn(B, 3, 1024, 1024, generator=gen).pin_memory()
    dev_tiles = host_tiles.to(dev)
    sizes = torch.tensor([[1024, 1024]] * B, device=dev)
    enc_eng, dec_eng = model.image_encoder.engine(), model.mask_decoder.transformer.engine()
    buf = pp.DetectionBuffer(B, Q, dev)

    def step(tiles):
        with torch.no_grad():
            out = model(NestedTensor(tiles, None), None)
            packed, _labels, _query, counts = pp.postprocess_packed(out["pred_logits"], out["pred_boxes"], sizes, conf_thr, out=buf)
            pp.nms_packed(packed, counts, score_thr=nms_thr, iou_threshold=iou_thr, per_class=per_class, out=buf)
            return gather_buffer(buf)

    for _ in range(warmup):
        step(dev_tiles)
    l0 = enc_eng.launches + dec_eng.launches
    timer = profiler.start()
    sampler = ClockSampler(ctx.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    ms_eager = ctx.timed(steps, lambda: step(dev_tiles))
    profiler.stop()
    fam = timer.summary()
    launches = (enc_eng.launches + dec_eng.launches - l0) + 2 * steps  # + postprocess + batched NMS
    torch.cuda.synchronize()
    cand = int(buf.counts.sum().item())
    kept = int(buf.keep_cnt.sum().item())
    graphed, gather_mode = None, "eager NCCL all-gather after every step" if world > 1 else "single GPU (no exchange)"
    if use_graph:
        graphed = GraphedDetector(model, B, conf_thr=conf_thr, nms_score_thr=nms_thr, iou_thr=iou_thr, warmup=1,
                                  per_class=per_class, gather=world > 1)
        if world > 1:
            gather_mode = "ONE NCCL all-gather of the flat detection buffer, captured in the CUDA graph"
        graphed.static_in.copy_(dev_tiles)
        for _ in range(warmup):
            graphed.replay()
        ms = ctx.timed(steps, graphed.replay)
        torch.cuda.synchronize()
        assert int(graphed.buffer.counts.sum().item()) == cand and int(graphed.buffer.keep_cnt.sum().item()) == kept
    else:
        ms = ms_eager
    clocks = sampler.stop() if sampler else None
    res = {"model_type": model_type, "batch_per_gpu": B, "queries": Q, "n_gpus": world,
           "value": world * B * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps,
           "eager": {"value": world * B * steps / (ms_eager / 1e3), "unit": UNIT, "ms_per_step": ms_eager / steps},
           "gpu_launches": launches, "steps": steps, "warmup": warmup,
           "launch": "CUDA-graph replay of the same kernels" if graphed is not None else "eager (one launch per kernel)",
           "gather": gather_mode,
           "nms": {"score_thr": nms_thr, "iou_thr": iou_thr, "per_class": per_class,
                   "candidates_per_tile": cand / B, "kept_per_tile": kept / B}}
    if clocks is not None:
        res["clocks"] = clocks

    # ---- end to end: pinned host tiles -> H2D -> path -> D2H of the detection buffer, every step.  The tile feed is
    # double buffered (the H2D copy of step i+1 runs on a copy stream while step i computes, as a serving loop does);
    # every step still pays one full H2D of its inputs and one D2H of its detections inside the timed region, and the
    # caller synchronises on each step's detections.
    if want_e2e:
        copy_stream = torch.cuda.Stream(device=dev)
        dev_buf = [torch.empty_like(dev_tiles), torch.empty_like(dev_tiles)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"i": 0}
        host_det = torch.empty(buf.flat.numel()).pin_memory()

        def prefetch(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])  # the step that last read this buffer has finished
                dev_buf[slot].copy_(host_tiles, non_blocking=True)
                ready[slot].record(copy_stream)

        def e2e_step():
            slot = state["i"] & 1
            state["i"] += 1
            prefetch(slot ^ 1)  # next step's tiles
            torch.cuda.current_stream().wait_event(ready[slot])
            if graphed is not None:
                graphed.static_in.copy_(dev_buf[slot], non_blocking=True)  # device copy into the captured input
                consumed[slot].record()
                graphed.replay()
                src = graphed.buffer.flat
            else:
                step(dev_buf[slot])
                consumed[slot].record()
                src = buf.flat
            host_det.copy_(src, non_blocking=True)  # every rank reads back its own shard's detections
            torch.cuda.current_stream().synchronize()  # the caller consumes the detections of this step

        for s_ in (0, 1):
            consumed[s_].record()
        prefetch(0)
        e2e_step()
        ms_e2e = ctx.timed(steps, e2e_step)
        torch.cuda.synchronize()
        res["e2e"] = {"value": world * B * steps / (ms_e2e / 1e3), "unit": UNIT,
                      "h2d_bytes_per_step": host_tiles.numel() * 4, "d2h_bytes_per_step": host_det.numel() * 4,
                      "ms_per_step": ms_e2e / steps}
        del dev_buf, host_det

    # ---- roofline of the dominant kernel family (event-instrumented eager pass)
    peaks = load_peaks()
    name, r = max(fam.items(), key=lambda kv: kv[1]["ms"])
    if r["flop"] > 0:
        achieved = r["flop"] / (r["ms"] / 1e3) / 1e12
        roof = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops"],
                "traffic": NCU_TRAFFIC.get(name) if (model_type, B) == ("vit_b", 32) else None,
                "traffic_note": "DRAM read+write bytes of the qkv-shaped launch (M=131072, N=2304, K=768; algorithmic 809 MB) "
                                "from one ncu --set full capture (profiles/); a constant, not measurable inside the run",
                "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside a long step)",
                "share_of_step": r["ms"] / ms_eager, "launches_per_step": r["launches"] / steps,
                "avg_launch_ms": r["ms"] / r["launches"],
                "measured_in": "event-instrumented eager pass over the same K steps (one CUDA-event pair per launch)"}
    else:
        achieved = r["byte"] / (r["ms"] / 1e3) / 1e9
        roof = {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peaks["gbs"], "unit": "GB/s",
                "frac": achieved / peaks["gbs"], "traffic": None, "peak_source": peaks["src"],
                "share_of_step": r["ms"] / ms_eager}
    res["roofline"] = roof
    res["breakdown"] = {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps,
                            "tflops": (v["flop"] / (v["ms"] / 1e3) / 1e12) if v["flop"] else None,
                            "gbs": (v["byte"] / (v["ms"] / 1e3) / 1e9) if v["byte"] else None}
                        for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    gf = GFLOP_PER_TILE.get(model_type)
    if gf:
        res["model_tflops"] = res["value"] * gf / 1e3
        res["model_frac_of_bf16_peak"] = res["value"] / world * gf / 1e3 / peaks["tflops"]
    # release this configuration's model, workspaces and graph before the next one
    del graphed, model, enc_eng, dec_eng, dev_tiles, host_tiles, buf
    gc.collect()
    torch.cuda.empty_cache()
    return res


def measure_nms_10k(ctx: Ctx, cpu_too: bool):
    """Per-class NMS over the 10 k-box dense-herd problem (wm_nms: rank + bitmask + reduce kernels), timed with CUDA events
    after warm-up; bytes = the n x ceil(n/64) x 8 B suppression mask written and re-read by the reduce."""
    torch = ctx.torch
    from wildlifemapper_b200.ops import ops
    boxes, scores, labels = dense_herd_boxes()
    n = boxes.shape[0]
    dev = ctx.dev
    b, s, l = torch.from_numpy(boxes).to(dev), torch.from_numpy(scores).to(dev), torch.from_numpy(labels).to(dev)
    nb = (n + 63) // 64
    keep = torch.empty(n, device=dev, dtype=torch.int64)
    num = torch.zeros(1, device=dev, dtype=torch.int32)
    order_ws = torch.empty(n, device=dev, dtype=torch.int32)
    mask_ws = torch.empty(n * nb, device=dev, dtype=torch.int64)

    def run():
        ops.nms(b, s, l, 0.4, order_ws, mask_ws, keep, num)

    for _ in range(3):
        run()
    ms = ctx.timed(10, run) / 10
    kept = int(num.item())
    mask_bytes = n * nb * 8
    out = {"n_boxes": n, "per_class": True, "iou_thr": 0.4, "ms": ms, "kept": kept, "mask_bytes": mask_bytes,
           "mask_gbs": 2 * mask_bytes / (ms / 1e3) / 1e9,
           "note": "rank (exact O(n^2) stable order) + bitmask IoU + serial reduce; mask written once and read once"}
    if cpu_too:
        try:
            import torchvision
            tb, ts, tl = torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(labels)

            def cpu_nms():  # per-class loop of torchvision.ops.nms (the oracle definition)
                keeps = []
                for c in range(7):
                    idx = torch.nonzero(tl == c).flatten()
                    keeps.append(idx[torchvision.ops.nms(tb[idx], ts[idx], 0.4)])
                k = torch.cat(keeps)
                return k[torch.argsort(ts[k], descending=True, stable=True)]

            cpu_nms()
            t0 = time.perf_counter()
            kk = cpu_nms()
            out["cpu_torchvision_ms"] = (time.perf_counter() - t0) * 1e3
            out["cpu_kept"] = int(kk.numel())
            out["matches_cpu"] = bool(kk.numel() == kept and torch.equal(kk, keep[:kept].cpu()))
        except Exception as e:
            out["cpu_torchvision_ms"] = None
            out["cpu_error"] = f"{type(e).__name__}: {e}"
    return out


def run_b200(args):
    sys.path.insert(0, os.path.join(ROOT, "wildlifemapper_b200"))  # the drop-in `segment_anything` package
    ctx = Ctx()
    torch = ctx.torch
    world, rank = ctx.world, ctx.rank
    B, Q = args.batch, args.queries
    head = measure_config(ctx, args.model, B, Q, args.steps, args.warmup, use_graph=not args.no_graph, sample_clocks=True)
    extras = {}
    dense = None
    if not args.no_extras and args.model == "vit_b":
        xs, xw = 3, 3  # short passes: W >= 3 warm-up steps, 3 timed steps
        try:
            extras["vit_l"] = measure_config(ctx, "vit_l", 32, 51, xs, xw)
            extras["vit_h"] = measure_config(ctx, "vit_h", 64, 51, xs, xw)
            dense = measure_config(ctx, "vit_b", 32, 900, xs, xw, per_class=True)
            dense["nms_10k"] = measure_nms_10k(ctx, cpu_too=(rank == 0 and world == 1))
        except Exception as e:  # the headline must survive a failure in the extra passes
            extras["error"] = f"{type(e).__name__}: {e}"
    if rank != 0:
        if world > 1:
            ctx.dist.destroy_process_group()
        return
    if args.breakdown:
        with open(args.breakdown, "w") as f:
            json.dump({"ms_per_step": head["eager"]["ms_per_step"], "batch": B, "model": args.model,
                       "families": head["breakdown"]}, f, indent=1)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_tiles_per_sec(args.model, 2, 1, Q)
        cpu = {"value": r["tps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
               "sample": f"2 timed + 1 warm-up single-tile forwards ({args.model}, fp32, {r['threads']} threads): {r['note']}"}

    def slim(c):
        if c is None:
            return None
        c = dict(c)
        c["breakdown_ms_per_step"] = {k: round(v["ms_per_step"], 3) for k, v in c.pop("breakdown").items()}
        return c

    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} detector (fft + encoder + decoder + PostProcess + NMS), batch {B} "
                                   f"synthetic 1024x1024 tiles per GPU, {Q} queries", "model_type": args.model,
                       "batch_per_gpu": B, "parallelism": f"tile-sharded x{world}, {head['gather']}",
                       "l2": "inputs and activations per step exceed L2 (no flush needed)", "launch": head["launch"],
                       "nms": "score threshold lowered from the reference's 0.5 to 0.05 so that every PostProcess row of the "
                              "random-init model is an NMS candidate (non-empty NMS work); IoU 0.4, class-agnostic",
                       "hfc_precision": os.environ.get("WM_HFC_PRECISION", "split") +
                                        " (MedSAM.fft low-pass on hi + lo bf16 operands: 3x the DFT-operator GEMM work, the "
                                        "reference's fp32 FFT to 1e-5; 'bf16' = round-1 single operands)"},
            "eager": head["eager"], "e2e": head.get("e2e"), "gpu_launches": head["gpu_launches"], "clocks": head.get("clocks"),
            "roofline": head["roofline"], "cpu_baseline": cpu, "nms": head["nms"],
            "model_tflops": head.get("model_tflops"), "model_frac_of_bf16_peak": head.get("model_frac_of_bf16_peak"),
            "breakdown_ms_per_step": {k: round(v["ms_per_step"], 3) for k, v in head["breakdown"].items()},
            "configs": {k: (slim(v) if isinstance(v, dict) else v) for k, v in extras.items()},
            "dense_herd": slim(dense)}
    emit(line)
    if world > 1:
        ctx.dist.destroy_process_group()


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
