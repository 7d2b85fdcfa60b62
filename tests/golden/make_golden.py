"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``segment_anything`` from /root/reference/wildlifemapper, loads the deterministic
synthetic weights from ``oracle/weights.py`` into the reference modules, runs the reference forward /
PostProcess / torchvision NMS on seeded inputs and stores small fixtures:

  golden_model_<cfg>.npz  logits, boxes, and per-stage samples (fixed pseudo-random positions)
  golden_post.npz         PostProcess outputs for seeded logits/boxes
  golden_nms.npz          torchvision.ops.nms keep lists (class-agnostic and per-class)
  golden_frontend.npz     ToTensor + Normalize + nested_tensor_from_tensor_list on seeded uint8 crops (hash + samples),
                          convert_to_xywh on seeded boxes
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/wildlifemapper"

from oracle import post as opost  # noqa: E402
from oracle.frontend import FRONTEND_CASES, RESIZE_CASES, frontend_image  # noqa: E402
from oracle.weights import MODEL_CONFIGS, make_state_dict, make_tiles  # noqa: E402

N_SAMPLES = 256


def sample_positions(numel: int, key: str) -> np.ndarray:
    import zlib
    rng = np.random.default_rng(zlib.crc32(key.encode()))
    return rng.integers(0, numel, N_SAMPLES)


def tap_summary(name: str, t: torch.Tensor) -> dict:
    f = t.detach().float().contiguous().view(-1)
    pos = sample_positions(f.numel(), name)
    return {f"{name}.samples": f[torch.from_numpy(pos)].numpy(),
            f"{name}.stats": np.array([f.mean().item(), f.abs().mean().item(), f.abs().max().item()], np.float64)}


def build_reference(model_type: str, num_queries: int):
    sys.path.insert(0, REF)
    from segment_anything.modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, TwoWayTransformer
    from segment_anything.network import MedSAM
    from functools import partial
    D, depth, heads, glob = MODEL_CONFIGS[model_type]
    enc = ImageEncoderViT(depth=depth, embed_dim=D, img_size=1024, mlp_ratio=4,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=heads, patch_size=16,
                          qkv_bias=True, use_rel_pos=True, global_attn_indexes=list(glob), window_size=14,
                          out_chans=256)
    pe = PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16)
    dec = MaskDecoder(num_multimask_outputs=num_queries - 1,
                      transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                      transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256)
    model = MedSAM(image_encoder=enc, mask_decoder=dec, prompt_encoder=pe).eval()
    return model


def golden_model(model_type: str, batch: int, num_queries: int, tag: str) -> None:
    sd = make_state_dict(model_type, seed=0, num_queries=num_queries)
    model = build_reference(model_type, num_queries)
    ref_keys = list(model.state_dict().keys())
    assert ref_keys == list(sd.keys()), "state_dict key contract drifted"
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    from segment_anything.utils.misc import NestedTensor
    tiles = make_tiles(batch, seed=2)
    taps = {}

    def hook(name):
        def fn(_m, _i, o):
            taps[name] = o
        return fn

    enc = model.image_encoder
    enc.patch_embed.register_forward_hook(hook("patch"))
    enc.hfc_embed.register_forward_hook(hook("hfc_tok"))
    enc.hfc_attn.register_forward_hook(hook("hfc_attn_out"))
    for i, blk in enumerate(enc.blocks):
        blk.register_forward_hook(hook(f"block{i}"))
    enc.register_forward_hook(hook("features"))
    model.mask_decoder.transformer.register_forward_hook(hook("transformer"))
    with torch.no_grad():
        x_hfc = model.fft(NestedTensor(tiles, None))
        out = model(NestedTensor(tiles, None), np.array([[0, 0, 1024, 1024]] * batch))
    res = {"pred_logits": out["pred_logits"].numpy(), "pred_boxes": out["pred_boxes"].numpy()}
    res.update(tap_summary("x_hfc", x_hfc))
    for k, v in taps.items():
        if k == "transformer":
            res.update(tap_summary("hs", v[0]))
        else:
            res.update(tap_summary(k, v))
    # reference PostProcess on the reference outputs
    sys.path.insert(0, REF)
    from segment_anything.build_sam import PostProcess
    ts = torch.tensor([[1024, 1024]] * batch)
    pp = PostProcess(0.05)(out, ts)
    for i, r in enumerate(pp):
        res[f"pp{i}.scores"] = r["scores"].numpy()
        res[f"pp{i}.labels"] = r["labels"].numpy()
        res[f"pp{i}.boxes"] = r["boxes"].numpy()
    res["meta"] = np.array([batch, num_queries])
    np.savez_compressed(os.path.join(HERE, f"golden_model_{tag}.npz"), **res)
    print("wrote", tag, {k: v.shape for k, v in res.items() if not k.endswith(("samples", "stats"))})


def golden_post() -> None:
    sys.path.insert(0, REF)
    from segment_anything.build_sam import PostProcess
    g = torch.Generator().manual_seed(11)
    res = {}
    for case, (B, Q, spread) in enumerate(((3, 51, 2.0), (2, 900, 4.0), (2, 51, 0.01))):
        logits = torch.randn(B, Q, 8, generator=g) * spread
        if case == 2:
            logits[..., -1] += 8.0  # everything below threshold in image 0 -> empty result branch
            logits[1, 5, 2] += 12.0
        boxes = torch.rand(B, Q, 4, generator=g) * 0.5 + 0.1
        ts = torch.tensor([[1024, 1024], [768, 1000], [3648, 5472]][:B])
        out = PostProcess(0.05)({"pred_logits": logits, "pred_boxes": boxes}, ts)
        res[f"c{case}.logits"] = logits.numpy()
        res[f"c{case}.boxes"] = boxes.numpy()
        res[f"c{case}.sizes"] = ts.numpy()
        for i, r in enumerate(out):
            res[f"c{case}.{i}.scores"] = r["scores"].numpy().astype(np.float32)
            res[f"c{case}.{i}.labels"] = r["labels"].numpy().astype(np.int64)
            res[f"c{case}.{i}.boxes"] = r["boxes"].numpy().astype(np.float32).reshape(-1, 4)
    np.savez_compressed(os.path.join(HERE, "golden_post.npz"), **res)
    print("wrote golden_post")


def golden_nms() -> None:
    import torchvision
    res = {"torchvision": np.array(torchvision.__version__)}
    for tag, n, dup in (("n2000", 2000, False), ("n2000dup", 2000, True), ("n10000", 10000, False)):
        b, s, l = opost.make_nms_problem(n, seed=3, dup_scores=dup)
        tb, ts_ = torch.from_numpy(b), torch.from_numpy(s)
        keep = torchvision.ops.nms(tb, ts_, 0.4).numpy()
        kc = []
        for c in range(7):  # per-class loop == _batched_nms_vanilla semantics
            idx = np.nonzero(l == c)[0]
            kc.append(idx[torchvision.ops.nms(tb[idx], ts_[idx], 0.4).numpy()])
        kc = np.sort(np.concatenate(kc))  # ties: lower original index first (as _batched_nms_vanilla's where())
        kc = kc[np.argsort(-s[kc], kind="stable")]
        res[f"{tag}.keep"] = keep.astype(np.int32)
        res[f"{tag}.keep_per_class"] = kc.astype(np.int32)
    # exact-tie probes: IoU == 2/5 vs threshold 0.4 in double
    tie_boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 4], [0, 0, 4, 10], [20, 20, 30, 30], [20, 20, 30, 24]], np.float32)
    tie_scores = np.array([0.9, 0.8, 0.7, 0.7, 0.7], np.float32)
    res["tie.boxes"], res["tie.scores"] = tie_boxes, tie_scores
    res["tie.keep"] = torchvision.ops.nms(torch.from_numpy(tie_boxes), torch.from_numpy(tie_scores), 0.4).numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "golden_nms.npz"), **res)
    print("wrote golden_nms", {k: (v.shape if hasattr(v, "shape") else v) for k, v in res.items()})


def golden_frontend() -> None:
    import hashlib
    import re
    from PIL import Image
    sys.path.insert(0, REF)
    import segment_anything.utils.augmentation as T
    from segment_anything.utils.misc import nested_tensor_from_tensor_list
    norm = T.Compose([T.ToTensor(), T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])  # dataloader_coco.py:286-290
    res = {}
    for tag, hw, (y0, x0), (ch, cw) in FRONTEND_CASES:
        img = frontend_image(tag, hw)
        crop = np.ascontiguousarray(img[y0:y0 + ch, x0:x0 + cw])  # the tile file the reference's loader would open
        t, _ = norm(Image.fromarray(crop), None)
        nt = nested_tensor_from_tensor_list([t])
        out = nt.tensors[0].numpy()
        assert out.shape == (3, 1024, 1024)
        res[f"{tag}.sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest())
        res[f"{tag}.samples"] = out.reshape(-1)[sample_positions(out.size, tag)]
        res[f"{tag}.sum"] = np.array([np.abs(out.astype(np.float64)).sum()])
    # RandomResize([768], max_size=768) (dataloader_coco.py:277,288 -> augmentation.py:77-107 -> torchvision F.resize -> PIL)
    for tag, hw, size, max_size in RESIZE_CASES:
        img = frontend_image(tag, hw)
        out, _ = T.RandomResize([size], max_size=max_size)(Image.fromarray(img), None)
        arr = np.asarray(out)
        res[f"{tag}.shape"] = np.array(arr.shape[:2])
        res[f"{tag}.sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest())
        res[f"{tag}.samples"] = arr.reshape(-1)[sample_positions(arr.size, tag)]
    # the loader pipelines of dataloader_coco.py:275-292 on an image + target (boxes xyxy, area, center), incl. the train flip
    import random
    for tag, hw in (("pipe_sq", (1024, 1024)), ("pipe_wide", (600, 900))):
        img = frontend_image(tag, hw)
        g = torch.Generator().manual_seed(11)
        xy = torch.rand(9, 2, generator=g) * torch.tensor([hw[1] - 80.0, hw[0] - 80.0])
        wh = torch.rand(9, 2, generator=g) * 60 + 8
        tgt = {"boxes": torch.cat([xy, xy + wh], 1), "area": wh[:, 0] * wh[:, 1], "center": xy + wh / 2,
               "labels": torch.arange(9) % 6 + 1, "orig_size": torch.as_tensor([hw[0], hw[1]]), "size": torch.as_tensor([hw[0], hw[1]])}
        for mode, tf in (("val", [T.RandomResize([768], max_size=768), T.ToTensor(), T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])]),
                         ("train", [T.RandomResize([768], max_size=768), T.ToTensor(), T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
                                    T.FlipLR(fliplr=1.0)])):
            random.seed(0)
            im, tg = T.Compose(tf)(Image.fromarray(img), {k: v.clone() for k, v in tgt.items()})
            a = im.numpy()
            res[f"{tag}.{mode}.sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest())
            res[f"{tag}.{mode}.shape"] = np.array(a.shape)
            for k in ("boxes", "area", "center", "size"):
                res[f"{tag}.{mode}.{k}"] = tg[k].numpy()
    # convert_to_xywh lives in inference.py, which does not import here (pycocotools): run its source text
    src = open(os.path.join(REF, "inference.py")).read()
    m = re.search(r"^def convert_to_xywh\(boxes\):\n(?:[ \t]+.*\n)+", src, re.M)
    ns = {"torch": torch}
    exec(m.group(0), ns)
    g = torch.Generator().manual_seed(5)
    xy = torch.rand(64, 2, generator=g) * 5000
    wh = torch.rand(64, 2, generator=g) * 60 + 1
    boxes = torch.cat([xy, xy + wh], 1)
    res["xywh.boxes"] = boxes.numpy()
    res["xywh.out"] = ns["convert_to_xywh"](boxes).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_frontend.npz"), **res)
    print("wrote golden_frontend", {k: (v.shape if v.ndim else str(v)[:16]) for k, v in res.items()})


def golden_criterion() -> None:
    """The reference's own HungarianMatcher + SetCriterion (build through sam_model_registry's args path) on seeded cases."""
    sys.path.insert(0, REF)
    from segment_anything.build_sam import SetCriterion
    from segment_anything.modeling.matcher import HungarianMatcher
    from oracle.criterion import CRITERION_CASES, make_case
    matcher = HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)  # train.py defaults: set_cost_class / bbox / giou
    crit = SetCriterion(7, matcher=matcher, weight_dict={"loss_ce": 3, "loss_bbox": 5, "loss_giou": 2}, eos_coef=0.1,
                        losses=["labels", "boxes", "cardinality"]).eval()
    res = {}
    for tag, B, Q, sizes in CRITERION_CASES:
        logits, boxes, targets = make_case(tag, B, Q, sizes)
        out = {"pred_logits": torch.from_numpy(logits), "pred_boxes": torch.from_numpy(boxes)}
        tg = [{"labels": torch.from_numpy(t["labels"]), "boxes": torch.from_numpy(t["boxes"])} for t in targets]
        with torch.no_grad():
            idx = matcher(out, tg)
            losses = crit(out, tg)
        for i, (a, b) in enumerate(idx):
            res[f"{tag}.idx{i}.src"], res[f"{tag}.idx{i}.tgt"] = a.numpy(), b.numpy()
        for k, v in losses.items():
            res[f"{tag}.{k}"] = np.array(float(v), np.float64)
        if sum(sizes):
            ids = torch.cat([t["labels"] for t in tg]); tb = torch.cat([t["boxes"] for t in tg])
            prob = out["pred_logits"].flatten(0, 1).softmax(-1)
            from segment_anything.utils.box_ops import box_cxcywh_to_xyxy, generalized_box_iou
            C = 5 * torch.cdist(out["pred_boxes"].flatten(0, 1), tb, p=1) + 1 * (-prob[:, ids]) + 2 * (-generalized_box_iou(
                box_cxcywh_to_xyxy(out["pred_boxes"].flatten(0, 1)), box_cxcywh_to_xyxy(tb)))
            c = C.numpy().reshape(-1)
            res[f"{tag}.cost.samples"] = c[sample_positions(c.size, tag + ".cost")]
    np.savez_compressed(os.path.join(HERE, "golden_criterion.npz"), **res)
    print("wrote golden_criterion", sorted(k for k in res if "idx" not in k))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 8)
    which = sys.argv[1:] or ["post", "nms", "frontend", "criterion", "vit_t", "vit_t900", "vit_b"]
    if "frontend" in which:
        golden_frontend()
    if "criterion" in which:
        golden_criterion()
    if "post" in which:
        golden_post()
    if "nms" in which:
        golden_nms()
    if "vit_t" in which:
        golden_model("vit_t", 2, 51, "vit_t")
    if "vit_t900" in which:
        golden_model("vit_t", 1, 900, "vit_t_q900")
    if "vit_b" in which:
        golden_model("vit_b", 1, 51, "vit_b")
