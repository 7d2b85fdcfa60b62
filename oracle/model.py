"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

fp32 CPU restatement of the WildlifeMapper tile-detection forward, written from the
distilled math in SURVEY.md App. A and checked against the imported reference by
``tests/golden/make_golden.py``.  Functional: it takes a reference-keyed ``state_dict``
(``oracle/weights.py``) and never instantiates reference modules.

Reference anchors (all under /root/reference/wildlifemapper/segment_anything/):
  hfc_highpass      network.py:36-57          (MedSAM.fft)
  patch/hfc embed   modeling/image_encoder.py:386-450, :123-131
  hfc_cross_attn    modeling/image_encoder.py:452-516 (incl. the raw reshape at :512)
  block             modeling/image_encoder.py:188-204, :246-262, :265-311, :314-383
  neck              modeling/image_encoder.py:105-121,136 ; modeling/common.py:31-43
  dense_pe          modeling/pos_encoder.py:24-33,50-70
  decoder           modeling/transformer.py:62-106,151-182,218-240 ; modeling/box_decoder.py:71-176
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .weights import GRID, MODEL_CONFIGS, WINDOW

Tensor = torch.Tensor


# --------------------------------------------------------------------------- E0
def hfc_highpass(img: Tensor, rate: float = 0.125) -> Tensor:
    """|ifft2(fftshift-masked fft2(gray))| ; network.py:36-57."""
    g = 0.2989 * img[:, 0:1] + 0.587 * img[:, 1:2] + 0.114 * img[:, 2:3]  # torchvision Grayscale
    w, h = g.shape[-2:]
    line = int((w * h * rate) ** 0.5 // 2)
    mask = torch.ones_like(g)
    mask[:, :, w // 2 - line:w // 2 + line, h // 2 - line:h // 2 + line] = 0
    f = torch.fft.fftshift(torch.fft.fft2(g, norm="forward"), dim=(-2, -1)) * mask
    inv = torch.fft.ifft2(torch.fft.ifftshift(f, dim=(-2, -1)), norm="forward").real
    return inv.abs()


def lowpass_operator(n: int = 1024, rate: float = 0.125, dtype=torch.float64):
    """Real/imag parts of L with lowpass(g) = Re(L g L^T); SURVEY.md App. A.1.

    keep[k] = 1 for DFT indices k in {-line .. line-1} (the zeroed box after fftshift).
    """
    line = int((n * n * rate) ** 0.5 // 2)
    k = torch.arange(-line, line, dtype=dtype)
    d = (torch.arange(n, dtype=dtype)[:, None] - torch.arange(n, dtype=dtype)[None, :])
    ang = 2 * math.pi * d[:, :, None] * k[None, None, :] / n
    return torch.cos(ang).sum(-1) / n, torch.sin(ang).sum(-1) / n


# --------------------------------------------------------------------------- blocks
def _rel_index(S: int) -> Tensor:
    # get_rel_pos with q_size == k_size: index q - k + (S - 1)   image_encoder.py:340-344
    c = torch.arange(S)
    return c[:, None] - c[None, :] + (S - 1)


def attention_relpos(x: Tensor, p: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """Attention.forward on [Bw, S, S, D] windows (S=14) or images (S=64); image_encoder.py:246-262."""
    Bw, S, _, D = x.shape
    hd = D // heads
    qkv = F.linear(x, p[pre + "qkv.weight"], p[pre + "qkv.bias"])
    qkv = qkv.reshape(Bw, S * S, 3, heads, hd).permute(2, 0, 3, 1, 4)  # [3,Bw,h,T,hd]
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)  # [Bw,h,T,T]
    idx = _rel_index(S)
    Rh = p[pre + "rel_pos_h"][idx]  # [S(q),S(k),hd]
    Rw = p[pre + "rel_pos_w"][idx]
    rq = q.reshape(Bw, heads, S, S, hd)  # unscaled q   image_encoder.py:256,375
    rel_h = torch.einsum("bnhwc,hkc->bnhwk", rq, Rh)
    rel_w = torch.einsum("bnhwc,wkc->bnhwk", rq, Rw)
    attn = (attn.view(Bw, heads, S, S, S, S) + rel_h[..., :, None] + rel_w[..., None, :]).view(Bw, heads, S * S, S * S)
    attn = attn.softmax(dim=-1)
    o = (attn @ v).permute(0, 2, 1, 3).reshape(Bw, S, S, D)
    return F.linear(o, p[pre + "proj.weight"], p[pre + "proj.bias"])


def block(x: Tensor, p: Dict[str, Tensor], pre: str, heads: int, window: int, chunk: int = 4) -> Tensor:
    """Block.forward; image_encoder.py:188-204. x: [B,64,64,D]."""
    B, H, W, D = x.shape
    xn = F.layer_norm(x, (D,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-6)
    if window > 0:
        pad = (window - H % window) % window
        xp = F.pad(xn, (0, 0, 0, pad, 0, pad))  # zero pad AFTER the norm   image_encoder.py:190-194,278-282
        Hp = H + pad
        nw = Hp // window
        win = xp.view(B, nw, window, nw, window, D).permute(0, 1, 3, 2, 4, 5).reshape(-1, window, window, D)
        a = attention_relpos(win, p, pre + "attn.", heads)  # pad keys participate
        a = a.view(B, nw, nw, window, window, D).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Hp, D)[:, :H, :W]
    else:
        # chunk the batch so the [h,4096,4096] score tensor stays small
        a = torch.cat([attention_relpos(xn[i:i + chunk], p, pre + "attn.", heads) for i in range(0, B, chunk)])
    x = x + a
    xn = F.layer_norm(x, (D,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-6)
    h = F.gelu(F.linear(xn, p[pre + "mlp.lin1.weight"], p[pre + "mlp.lin1.bias"]))  # exact erf GELU
    return x + F.linear(h, p[pre + "mlp.lin2.weight"], p[pre + "mlp.lin2.bias"])


# --------------------------------------------------------------------------- E3
def hfc_cross_attn(hfc_tok: Tensor, x: Tensor, p: Dict[str, Tensor], pre: str, nhead: int = 8) -> Tensor:
    """CrossAttentionHfcPatch.forward (eval: dropout = identity); image_encoder.py:486-516.

    hfc_tok [B,64,64,1024], x [B,64,64,D] -> [B,64,64,D].
    """
    B, H, W, C = hfc_tok.shape
    N = H * W
    D = x.shape[-1]
    pos = p[pre + "pos_embed"].reshape(C, N).t()  # [N,C]
    hk = F.linear(hfc_tok.reshape(B, N, C), p[pre + "proj_hfc.weight"].reshape(C, C), p[pre + "proj_hfc.bias"]) + pos
    pq = F.linear(x.reshape(B, N, D), p[pre + "proj_patch.weight"].reshape(C, D), p[pre + "proj_patch.bias"])
    Wi, bi = p[pre + "cross_attn.in_proj_weight"], p[pre + "cross_attn.in_proj_bias"]
    hd = C // nhead
    q = F.linear(pq, Wi[:C], bi[:C]) * (1.0 / math.sqrt(hd))  # scaled after bias (torch MHA slow path)
    k = F.linear(hk, Wi[C:2 * C], bi[C:2 * C])
    v = F.linear(hk, Wi[2 * C:], bi[2 * C:])
    sp = lambda t: t.view(B, N, nhead, hd).transpose(1, 2)
    o = torch.cat([(sp(q)[i:i + 2] @ sp(k)[i:i + 2].transpose(-2, -1)).softmax(-1) @ sp(v)[i:i + 2] for i in range(0, B, 2)])
    o = o.transpose(1, 2).reshape(B, N, C)
    src2 = F.linear(o, p[pre + "cross_attn.out_proj.weight"], p[pre + "cross_attn.out_proj.bias"])
    z = F.layer_norm(pq + src2, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
    f = F.linear(F.relu(F.linear(z, p[pre + "linear1.weight"], p[pre + "linear1.bias"])),
                 p[pre + "linear2.weight"], p[pre + "linear2.bias"])
    z = F.layer_norm(f + z, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
    # image_encoder.py:512 -- RAW reinterpret of the token-major [B,N,C] buffer as [B,C,H,W]
    y = z.contiguous().reshape(B, C, N)
    out = torch.einsum("oc,bcp->bpo", p[pre + "proj_back.weight"].reshape(D, C), y) + p[pre + "proj_back.bias"]
    return out.reshape(B, H, W, D)


# --------------------------------------------------------------------------- encoder
def _ln2d(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[:, None, None]


def image_encoder(sd: Dict[str, Tensor], model_type: str, img: Tensor, x_hfc: Tensor,
                  taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """ImageEncoderViT.forward; image_encoder.py:123-138. Returns [B,256,64,64]."""
    D, depth, heads, glob = MODEL_CONFIGS[model_type]
    e = "image_encoder."
    x = F.conv2d(img, sd[e + "patch_embed.proj.weight"], sd[e + "patch_embed.proj.bias"], stride=16).permute(0, 2, 3, 1)
    x = x + sd[e + "pos_embed"]
    ht = F.conv2d(x_hfc, sd[e + "hfc_embed.proj.weight"], sd[e + "hfc_embed.proj.bias"], stride=16).permute(0, 2, 3, 1)
    if taps is not None:
        taps["patch_pos"] = x
        taps["hfc_tok"] = ht
    xh = hfc_cross_attn(ht, x, sd, e + "hfc_attn.")
    x = xh + x
    if taps is not None:
        taps["after_hfc"] = x
    for i in range(depth):
        x = block(x, sd, f"{e}blocks.{i}.", heads, 0 if i in glob else WINDOW)
        if taps is not None:
            taps[f"block{i}"] = x
    t = F.conv2d(x.permute(0, 3, 1, 2), sd[e + "neck.0.weight"])
    t = _ln2d(t, sd[e + "neck.1.weight"], sd[e + "neck.1.bias"])
    t = F.conv2d(t, sd[e + "neck.2.weight"], padding=1)
    return _ln2d(t, sd[e + "neck.3.weight"], sd[e + "neck.3.bias"])


# --------------------------------------------------------------------------- decoder
def dense_pe(G: Tensor, size: int = GRID) -> Tensor:
    """PromptEncoder.get_dense_pe; pos_encoder.py:24-33,50-70. Returns [1,256,size,size]."""
    g = (torch.arange(size, dtype=torch.float32, device=G.device) + 0.5) / size
    xy = torch.stack([g[None, :].expand(size, size), g[:, None].expand(size, size)], dim=-1)  # (x, y)
    c = 2 * math.pi * ((2 * xy - 1) @ G)
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1).permute(2, 0, 1).unsqueeze(0)


def _dec_attn(q: Tensor, k: Tensor, v: Tensor, p: Dict[str, Tensor], pre: str, heads: int = 8) -> Tensor:
    """decoder Attention.forward; transformer.py:218-240."""
    q = F.linear(q, p[pre + "q_proj.weight"], p[pre + "q_proj.bias"])
    k = F.linear(k, p[pre + "k_proj.weight"], p[pre + "k_proj.bias"])
    v = F.linear(v, p[pre + "v_proj.weight"], p[pre + "v_proj.bias"])
    B, Tq, C = q.shape
    ch = C // heads
    sp = lambda t: t.reshape(t.shape[0], t.shape[1], heads, ch).transpose(1, 2)
    a = torch.softmax((sp(q) @ sp(k).transpose(-2, -1)) / math.sqrt(ch), dim=-1)
    o = (a @ sp(v)).transpose(1, 2).reshape(B, Tq, C)
    return F.linear(o, p[pre + "out_proj.weight"], p[pre + "out_proj.bias"])


def _ln(x: Tensor, p: Dict[str, Tensor], pre: str, eps: float = 1e-5) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), p[pre + "weight"], p[pre + "bias"], eps)


def box_decoder(sd: Dict[str, Tensor], feats: Tensor, taps: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """MaskDecoder.forward with the TwoWayTransformer; box_decoder.py:71-149, transformer.py:62-106."""
    B = feats.shape[0]
    pe = dense_pe(sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]).flatten(2).permute(0, 2, 1)
    keys = feats.flatten(2).permute(0, 2, 1)  # [B,4096,256]
    tokens = sd["mask_decoder.mask_tokens.weight"].unsqueeze(0).expand(B, -1, -1)
    queries = tokens
    t = "mask_decoder.transformer."
    for i in range(2):
        L = f"{t}layers.{i}."
        if i == 0:  # skip_first_layer_pe: output REPLACES the queries   transformer.py:155-156
            queries = _dec_attn(queries, queries, queries, sd, L + "self_attn.")
        else:
            q = queries + tokens
            queries = queries + _dec_attn(q, q, queries, sd, L + "self_attn.")
        queries = _ln(queries, sd, L + "norm1.")
        queries = queries + _dec_attn(queries + tokens, keys + pe, keys, sd, L + "cross_attn_token_to_image.")
        queries = _ln(queries, sd, L + "norm2.")
        m = F.linear(F.relu(F.linear(queries, sd[L + "mlp.lin1.weight"], sd[L + "mlp.lin1.bias"])),
                     sd[L + "mlp.lin2.weight"], sd[L + "mlp.lin2.bias"])
        queries = _ln(queries + m, sd, L + "norm3.")
        keys = keys + _dec_attn(keys + pe, queries + tokens, queries, sd, L + "cross_attn_image_to_token.")
        keys = _ln(keys, sd, L + "norm4.")
    queries = queries + _dec_attn(queries + tokens, keys + pe, keys, sd, t + "final_attn_token_to_image.")
    hs = _ln(queries, sd, t + "norm_final_attn.")
    if taps is not None:
        taps["hs"] = hs

    def mlp3(x: Tensor, pre: str) -> Tensor:
        for i in range(3):
            x = F.linear(x, sd[f"{pre}layers.{i}.weight"], sd[f"{pre}layers.{i}.bias"])
            if i < 2:
                x = F.relu(x)
        return x

    return {"pred_logits": mlp3(hs, "mask_decoder.class_embed."),
            "pred_boxes": mlp3(hs, "mask_decoder.bbox_embed.").sigmoid()}


@torch.no_grad()
def forward(sd: Dict[str, Tensor], model_type: str, tiles: Tensor,
            taps: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """MedSAM.forward; network.py:59-87. tiles [B,3,1024,1024] fp32."""
    x_hfc = hfc_highpass(tiles)
    feats = image_encoder(sd, model_type, tiles, x_hfc, taps)
    if taps is not None:
        taps["x_hfc"] = x_hfc
        taps["features"] = feats
    return box_decoder(sd, feats, taps)
