"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Deterministic synthetic weights for the WildlifeMapper detector, keyed exactly like the
reference ``state_dict`` (reference modules
``segment_anything/modeling/image_encoder.py``, ``transformer.py``, ``box_decoder.py``,
``pos_encoder.py``).  Every tensor is drawn from its own generator seeded by
(seed, crc32(key)), so the values do not depend on construction order and the same
call gives the same bytes in the build container and on the GPU box.

Zero-initialised reference parameters (pos_embed, rel_pos_h/w, hfc_attn.pos_embed, MHA
biases) are deliberately given non-zero values: a parity test on the reference's
default init would never exercise those code paths.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

MODEL_CONFIGS = {
    # name: (embed_dim, depth, heads, global_attn_indexes)   reference build_sam.py:19-52
    "vit_b": (768, 12, 12, (2, 5, 8, 11)),
    "vit_l": (1024, 24, 16, (5, 11, 17, 23)),
    "vit_h": (1280, 32, 16, (7, 15, 23, 31)),
    # tiny config for fast CPU tests (not a reference model; same structure)
    "vit_t": (128, 2, 2, (1,)),
    # same, with ViT-H's head dim 80 (8 heads x 80): exercises the head-dim-80 attention kernels
    "vit_t80": (640, 2, 8, (1,)),
}

HFC_DIM = 1024
OUT_CHANS = 256
WINDOW = 14
GRID = 64
NUM_LOGITS = 8  # 7 classes + no-object   reference box_decoder.py:50,68


def state_dict_spec(model_type: str, num_queries: int = 51) -> "OrderedDict[str, Tuple[int, ...]]":
    """Key -> shape, in the reference's registration order."""
    D, depth, heads, glob = MODEL_CONFIGS[model_type]
    hd = D // heads
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    e = "image_encoder."
    s[e + "pos_embed"] = (1, GRID, GRID, D)
    s[e + "patch_embed.proj.weight"] = (D, 3, 16, 16)
    s[e + "patch_embed.proj.bias"] = (D,)
    s[e + "hfc_embed.proj.weight"] = (HFC_DIM, 1, 16, 16)
    s[e + "hfc_embed.proj.bias"] = (HFC_DIM,)
    a = e + "hfc_attn."
    s[a + "pos_embed"] = (1, HFC_DIM, GRID, GRID)
    s[a + "proj_hfc.weight"] = (HFC_DIM, HFC_DIM, 1, 1)
    s[a + "proj_hfc.bias"] = (HFC_DIM,)
    s[a + "proj_patch.weight"] = (HFC_DIM, D, 1, 1)
    s[a + "proj_patch.bias"] = (HFC_DIM,)
    s[a + "cross_attn.in_proj_weight"] = (3 * HFC_DIM, HFC_DIM)
    s[a + "cross_attn.in_proj_bias"] = (3 * HFC_DIM,)
    s[a + "cross_attn.out_proj.weight"] = (HFC_DIM, HFC_DIM)
    s[a + "cross_attn.out_proj.bias"] = (HFC_DIM,)
    s[a + "linear1.weight"] = (HFC_DIM, HFC_DIM)
    s[a + "linear1.bias"] = (HFC_DIM,)
    s[a + "linear2.weight"] = (HFC_DIM, HFC_DIM)
    s[a + "linear2.bias"] = (HFC_DIM,)
    s[a + "norm1.weight"] = (HFC_DIM,)
    s[a + "norm1.bias"] = (HFC_DIM,)
    s[a + "norm2.weight"] = (HFC_DIM,)
    s[a + "norm2.bias"] = (HFC_DIM,)
    s[a + "proj_back.weight"] = (D, HFC_DIM, 1, 1)
    s[a + "proj_back.bias"] = (D,)
    for i in range(depth):
        b = f"{e}blocks.{i}."
        S = GRID if i in glob else WINDOW
        s[b + "norm1.weight"] = (D,)
        s[b + "norm1.bias"] = (D,)
        s[b + "attn.rel_pos_h"] = (2 * S - 1, hd)
        s[b + "attn.rel_pos_w"] = (2 * S - 1, hd)
        s[b + "attn.qkv.weight"] = (3 * D, D)
        s[b + "attn.qkv.bias"] = (3 * D,)
        s[b + "attn.proj.weight"] = (D, D)
        s[b + "attn.proj.bias"] = (D,)
        s[b + "norm2.weight"] = (D,)
        s[b + "norm2.bias"] = (D,)
        s[b + "mlp.lin1.weight"] = (4 * D, D)
        s[b + "mlp.lin1.bias"] = (4 * D,)
        s[b + "mlp.lin2.weight"] = (D, 4 * D)
        s[b + "mlp.lin2.bias"] = (D,)
    s[e + "neck.0.weight"] = (OUT_CHANS, D, 1, 1)
    s[e + "neck.1.weight"] = (OUT_CHANS,)
    s[e + "neck.1.bias"] = (OUT_CHANS,)
    s[e + "neck.2.weight"] = (OUT_CHANS, OUT_CHANS, 3, 3)
    s[e + "neck.3.weight"] = (OUT_CHANS,)
    s[e + "neck.3.bias"] = (OUT_CHANS,)
    t = "mask_decoder.transformer."

    def attn(prefix: str, internal: int) -> None:
        for n in ("q_proj", "k_proj", "v_proj"):
            s[f"{prefix}.{n}.weight"] = (internal, 256)
            s[f"{prefix}.{n}.bias"] = (internal,)
        s[f"{prefix}.out_proj.weight"] = (256, internal)
        s[f"{prefix}.out_proj.bias"] = (256,)

    for i in range(2):
        L = f"{t}layers.{i}."
        attn(L + "self_attn", 256)
        s[L + "norm1.weight"] = (256,)
        s[L + "norm1.bias"] = (256,)
        attn(L + "cross_attn_token_to_image", 128)
        s[L + "norm2.weight"] = (256,)
        s[L + "norm2.bias"] = (256,)
        s[L + "mlp.lin1.weight"] = (2048, 256)
        s[L + "mlp.lin1.bias"] = (2048,)
        s[L + "mlp.lin2.weight"] = (256, 2048)
        s[L + "mlp.lin2.bias"] = (256,)
        s[L + "norm3.weight"] = (256,)
        s[L + "norm3.bias"] = (256,)
        s[L + "norm4.weight"] = (256,)
        s[L + "norm4.bias"] = (256,)
        attn(L + "cross_attn_image_to_token", 128)
    attn(t + "final_attn_token_to_image", 128)
    s[t + "norm_final_attn.weight"] = (256,)
    s[t + "norm_final_attn.bias"] = (256,)
    s["mask_decoder.iou_token.weight"] = (1, 256)
    s["mask_decoder.mask_tokens.weight"] = (num_queries, 256)
    for i, (o, n) in enumerate(((256, 256), (256, 256), (NUM_LOGITS, 256))):
        s[f"mask_decoder.class_embed.layers.{i}.weight"] = (o, n)
        s[f"mask_decoder.class_embed.layers.{i}.bias"] = (o,)
    for i, (o, n) in enumerate(((256, 256), (256, 256), (4, 256))):
        s[f"mask_decoder.bbox_embed.layers.{i}.weight"] = (o, n)
        s[f"mask_decoder.bbox_embed.layers.{i}.bias"] = (o,)
    s["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"] = (2, 128)
    return s


def _fan_in(shape: Tuple[int, ...]) -> int:
    n = 1
    for d in shape[1:]:
        n *= d
    return max(n, 1)


def make_state_dict(model_type: str = "vit_b", seed: int = 0, num_queries: int = 51) -> Dict[str, torch.Tensor]:
    """fp32 CPU state_dict with reference key names and shapes."""
    out: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in state_dict_spec(model_type, num_queries).items():
        g = torch.Generator(device="cpu")
        g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2**63 - 1))
        leaf = key.rsplit(".", 1)[-1]
        is_norm = (".norm" in key or "neck.1." in key or "neck.3." in key) and "mlp" not in key
        if is_norm and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        elif "rel_pos" in key:
            t = 0.05 * torch.randn(shape, generator=g)
        elif leaf == "pos_embed":
            t = 0.05 * torch.randn(shape, generator=g)
        elif "gaussian_matrix" in key:
            t = torch.randn(shape, generator=g)
        elif "tokens.weight" in key or "iou_token" in key:
            t = torch.randn(shape, generator=g)
        elif leaf in ("bias", "in_proj_bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:  # linear / conv weights: N(0, 1/fan_in) keeps activations O(1)
            t = torch.randn(shape, generator=g) * (1.0 / _fan_in(shape)) ** 0.5
        out[key] = t.to(torch.float32).contiguous()
    return out


def make_tiles(batch: int, seed: int = 2, loader_like: bool = False) -> torch.Tensor:
    """Synthetic normalised RGB tiles [B,3,1024,1024] fp32.

    ``loader_like``: content only in the top-left 768x768, zeros elsewhere, like the
    reference loader's resize-then-pad (dataloader_coco.py:279, utils/misc.py:50-64).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.randn(batch, 3, 1024, 1024, generator=g, dtype=torch.float32)
    if loader_like:
        y = torch.zeros_like(x)
        y[:, :, :768, :768] = x[:, :, :768, :768]
        x = y
    return x
