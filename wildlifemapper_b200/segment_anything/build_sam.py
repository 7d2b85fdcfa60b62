"""Model factory, registry, SetCriterion and PostProcess with the reference names and signatures (reference
``segment_anything/build_sam.py``).  The criterion computes the forward values the reference's evaluation loop logs
(``inference.py:29-89``); its backward belongs to the training path, which is not implemented."""
from functools import partial

import torch
from torch import nn

from wildlifemapper_b200 import postprocess as _pp

from .modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, Sam, TwoWayTransformer
from .modeling.matcher import build_matcher


def build_sam_vit_h(checkpoint=None, args=None):
    return _build_sam(1280, 32, 16, [7, 15, 23, 31], checkpoint=checkpoint, args=args)


build_sam = build_sam_vit_h


def build_sam_vit_l(checkpoint=None, args=None):
    return _build_sam(1024, 24, 16, [5, 11, 17, 23], checkpoint=checkpoint, args=args)


def build_sam_vit_b(checkpoint=None, args=None):
    return _build_sam(768, 12, 12, [2, 5, 8, 11], checkpoint=checkpoint, args=args)


sam_model_registry = {
    "default": build_sam_vit_h,
    "vit_h": build_sam_vit_h,
    "vit_l": build_sam_vit_l,
    "vit_b": build_sam_vit_b,
}


class SetCriterion(nn.Module):
    """The DETR criterion of the reference (build_sam.py:62-210) for the EVALUATION loop: ``inference.py:29-89 evaluate()``
    calls ``criterion(outputs, targets)`` under ``torch.no_grad()`` for every batch.  Forward values only -- Hungarian
    matching (cost kernel + scipy, like the reference), then one sm_100a kernel for the weighted cross entropy, class
    error, cardinality error, L1 and GIoU box losses (``wm_set_criterion``).  Inputs that require grad raise: the
    backward is the training path, which this package does not implement."""

    def __init__(self, num_classes, matcher, weight_dict, eos_coef, losses):
        super().__init__()
        self.num_classes, self.matcher, self.eos_coef, self.losses = num_classes, matcher, eos_coef, list(losses)
        self.weight_dict = dict(weight_dict or {})
        empty_weight = torch.ones(self.num_classes + 1)
        empty_weight[-1] = self.eos_coef
        self.register_buffer("empty_weight", empty_weight)
        for name in self.losses:
            if name not in ("labels", "boxes", "cardinality"):
                raise NotImplementedError(f"do you really want to compute {name} loss?")

    def forward(self, outputs, targets):
        """-> {'loss_ce', 'class_error', 'loss_bbox', 'loss_giou', 'cardinality_error'}: 0-dim fp32 tensors on the device."""
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        if torch.is_grad_enabled() and (logits.requires_grad or boxes.requires_grad):
            raise RuntimeError("SetCriterion: inputs require grad, but wildlifemapper_b200 implements the evaluation path only "
                               "(no backward kernels). Run under torch.no_grad().")
        if self.matcher is None:
            raise RuntimeError("SetCriterion needs a matcher (sam_model_registry builds one from args.set_cost_*)")
        if "aux_outputs" in outputs:
            raise NotImplementedError("aux_outputs (aux_loss=True) are a training-only option")
        from wildlifemapper_b200.ops import ops
        from .utils.misc import get_world_size, is_dist_avail_and_initialized
        dev = logits.device
        B, Q, C1 = logits.shape
        indices = self.matcher({"pred_logits": logits, "pred_boxes": boxes}, targets)
        # matched pairs, image-major, in the matcher's order (== _get_src_permutation_idx, build_sam.py:152-156)
        rows = torch.cat([src + i * Q for i, (src, _) in enumerate(indices)]).to(device=dev, dtype=torch.int32)
        m_label = torch.cat([t["labels"].to(dev)[J.to(dev)] for t, (_, J) in zip(targets, indices)]).to(torch.int64).contiguous()
        m_box = torch.cat([t["boxes"].to(dev)[J.to(dev)] for t, (_, J) in zip(targets, indices)]).float().reshape(-1, 4).contiguous()
        tgt_len = torch.tensor([len(t["labels"]) for t in targets], device=dev, dtype=torch.int32)
        num_boxes = torch.as_tensor([sum(len(t["labels"]) for t in targets)], dtype=torch.float, device=dev)
        if is_dist_avail_and_initialized():
            torch.distributed.all_reduce(num_boxes)
        num_boxes = torch.clamp(num_boxes / get_world_size(), min=1).item()
        out5 = torch.empty(5, device=dev, dtype=torch.float32)
        tcls = torch.empty(B * Q, device=dev, dtype=torch.int32)
        ops.set_criterion(logits.detach().float().contiguous(), boxes.detach().float().contiguous(), rows, m_label, m_box, tgt_len,
                          self.empty_weight.to(device=dev, dtype=torch.float32), float(num_boxes), tcls, out5)
        res = {}
        if "labels" in self.losses:
            res["loss_ce"], res["class_error"] = out5[0], out5[1]
        if "boxes" in self.losses:
            res["loss_bbox"], res["loss_giou"] = out5[3], out5[4]
        if "cardinality" in self.losses:
            res["cardinality_error"] = out5[2]
        return res


class PostProcess(nn.Module):
    """Reference build_sam.py:212-258: softmax -> drop no-object -> max -> keep score > threshold ->
    cxcywh->xyxy -> scale by [s0, s1, s0, s1] (s = target_sizes row; the reference's h/w swap is reproduced)."""

    def __init__(self, confidence_threshold=0.05):
        super().__init__()
        self.confidence_threshold = confidence_threshold

    @torch.no_grad()
    def forward(self, outputs, target_sizes):
        out_logits, out_bbox = outputs["pred_logits"], outputs["pred_boxes"]
        assert len(out_logits) == len(target_sizes)
        assert target_sizes.shape[1] == 2
        packed, labels, _query, counts = _pp.postprocess_packed(out_logits, out_bbox, target_sizes,
                                                                self.confidence_threshold)
        results = []
        for i, n in enumerate(counts.tolist()):  # one D2H read: result lengths are data dependent
            if n == 0:  # same empty result as the reference (CPU tensors, build_sam.py:240-243)
                results.append({"scores": torch.tensor([]), "labels": torch.tensor([]),
                                "boxes": torch.tensor([]).reshape(0, 4)})
                continue
            results.append({"scores": packed[i, :n, 4], "labels": labels[i, :n], "boxes": packed[i, :n, :4]})
        return results


def _build_sam(encoder_embed_dim, encoder_depth, encoder_num_heads, encoder_global_attn_indexes, checkpoint=None,
               args=None):
    prompt_embed_dim = 256
    image_size = 1024
    vit_patch_size = 16
    image_embedding_size = image_size // vit_patch_size
    num_classes = 6 + 1
    # reference hard-codes 50 (+1) queries (build_sam.py:296); ``args.num_queries`` is the dense-herd knob
    num_queries = int(getattr(args, "num_queries", 51) or 51)
    sam = Sam(
        image_encoder=ImageEncoderViT(
            depth=encoder_depth, embed_dim=encoder_embed_dim, img_size=image_size, mlp_ratio=4,
            norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=encoder_num_heads, patch_size=vit_patch_size,
            qkv_bias=True, use_rel_pos=True, global_attn_indexes=encoder_global_attn_indexes, window_size=14,
            out_chans=prompt_embed_dim),
        prompt_encoder=PromptEncoder(embed_dim=prompt_embed_dim,
                                     image_embedding_size=(image_embedding_size, image_embedding_size),
                                     input_image_size=(image_size, image_size), mask_in_chans=16),
        mask_decoder=MaskDecoder(
            num_multimask_outputs=num_queries - 1,
            transformer=TwoWayTransformer(depth=2, embedding_dim=prompt_embed_dim, mlp_dim=2048, num_heads=8),
            transformer_dim=prompt_embed_dim, iou_head_depth=3, iou_head_hidden_dim=256),
        pixel_mean=[123.675, 116.28, 103.53],
        pixel_std=[58.395, 57.12, 57.375],
    )
    sam.eval()
    if checkpoint is not None:
        with open(checkpoint, "rb") as f:
            state_dict = torch.load(f, map_location="cpu")
        # keep only the transformer weights of the SAM mask decoder (reference build_sam.py:311-320)
        for k in [p for p in list(state_dict.keys()) if "mask_decoder" in p and "transformer" not in p]:
            del state_dict[k]
        sam.load_state_dict(state_dict, strict=False)
    weight_dict = {"loss_ce": 3, "loss_bbox": getattr(args, "bbox_loss_coef", 5),
                   "loss_giou": getattr(args, "giou_loss_coef", 2)}
    matcher = build_matcher(args) if args is not None and hasattr(args, "set_cost_class") else None
    criterion = SetCriterion(num_classes, matcher=matcher, weight_dict=weight_dict,
                             eos_coef=getattr(args, "eos_coef", 0.1), losses=["labels", "boxes", "cardinality"])
    device = getattr(args, "device", None)
    if device is not None:
        criterion.to(device)
    postprocessors = {"bbox": PostProcess(confidence_threshold=0.05)}
    return sam, criterion, postprocessors
