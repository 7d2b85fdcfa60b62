"""Kernel schedule of the tile-detection forward pass on one B200.

``EncoderEngine``  fft high-pass (network.py:36-57) + ImageEncoderViT.forward (image_encoder.py:123-138)
``DecoderEngine``  MaskDecoder.forward with the TwoWayTransformer (box_decoder.py:71-149, transformer.py:62-182)

The engines own only DERIVED device buffers (bf16 weight copies, fused layouts, rel-pos tables, the low-pass
DFT operator, the dense positional encoding) and per-batch workspaces; the fp32 master weights stay in the
``nn.Parameter``s of the drop-in modules (``segment_anything``), which hand their ``state_dict`` to
``prepare``.  Every compute step is one ``torch.ops.wm_b200.*`` call (hand-written sm_100a kernels); torch is
used for allocation only.

Data layout in HBM (M = B*4096 token rows, row = image-major, then y, then x):
  residual stream x   fp32 [M, D]            (kept in fp32: SURVEY.md section 0.7)
  GEMM operands       bf16 [M, *] row-major  (K contiguous -> K-major TMA/UMMA tiles)
  qkv                 bf16 [M, 3D]  columns [q | k | v] x [head] x [64]
  weights             bf16 [N, K] (nn.Linear layout, K contiguous)
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Tuple

import torch

from .profiler import credit, ops

ACT_NONE, ACT_GELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3
GRID = 64
NTOK = GRID * GRID
HFC = 1024


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous()


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def lowpass_operator_tables(n: int = 1024, rate: float = 0.125) -> Tuple[torch.Tensor, torch.Tensor]:
    """(Lr, Li) float64 with lowpass(g) = Lr g Lr^T - Li g Li^T  (MedSAM.fft, network.py:43-55).

    L[m, n] = f((m - n) mod N), f = ifft(keep), keep = 1 on DFT indices {-line .. line-1}.
    """
    line = int((n * n * rate) ** 0.5 // 2)
    keep = torch.zeros(n, dtype=torch.float64)
    k = torch.arange(-line, line) % n
    keep[k] = 1.0
    f = torch.fft.ifft(keep.to(torch.complex128))
    d = (torch.arange(n)[:, None] - torch.arange(n)[None, :]) % n
    return f.real[d].contiguous(), f.imag[d].contiguous()


class _Workspace:
    def __init__(self, device: torch.device):
        self.device = device
        self.buf: Dict[str, torch.Tensor] = {}

    def get(self, name: str, shape, dtype) -> torch.Tensor:
        t = self.buf.get(name)
        n = 1
        for s in shape:
            n *= s
        if t is None or t.dtype != dtype or t.numel() < n:
            t = torch.empty(n, device=self.device, dtype=dtype)
            self.buf[name] = t
        return t[:n].view(*shape)


def _gemm(a, w, bias=None, residual=None, res_mod=0, out_bf16=None, out_f32=None, act=ACT_NONE, bn=0):
    ops.gemm(a, w, bias, residual, res_mod if residual is not None else 0, out_bf16, out_f32, act, bn)


class EncoderEngine:
    """Derived weights + schedule for fft + ImageEncoderViT on one device."""

    def __init__(self, embed_dim: int, depth: int, num_heads: int, global_attn_indexes, device):
        if embed_dim % num_heads != 0 or embed_dim // num_heads not in (64, 80):
            raise NotImplementedError(
                f"head_dim {embed_dim / num_heads} is not supported by the tcgen05 attention kernels (64: ViT-B/L, 80: ViT-H)")
        self.D, self.depth, self.H = embed_dim, depth, num_heads
        self.hd = embed_dim // num_heads
        self.glob = tuple(global_attn_indexes)
        self.device = torch.device(device)
        self.ws = _Workspace(self.device)
        self.w: Dict[str, torch.Tensor] = {}
        self.launches = 0
        # generation counters of the workspaces whose views leave the engine (validity tokens of the aliases the drop-in
        # modules hang on API-visible tensors, see modeling/common.py::attach_twin): im2col rows / encoder features
        self.gen_rows = 0
        self.gen_feat = 0
        # low-pass products of MedSAM.fft: "split" = hi + lo bf16 terms of the image, the DFT operator and the intermediate
        # (3 tensor-core products per stage, ~2e-6 absolute, i.e. the reference's fp32 FFT); "bf16" = single bf16 operands
        # (1e-3 absolute: fine on noise-like tiles, 5 % of x_hfc on smooth 1/f^1.5 imagery).  WM_HFC_PRECISION overrides.
        self.hfc_precision = os.environ.get("WM_HFC_PRECISION", "split")
        if self.hfc_precision not in ("split", "bf16"):
            raise ValueError(f"WM_HFC_PRECISION must be 'split' or 'bf16', got {self.hfc_precision!r}")

    # ------------------------------------------------------------------ weight preparation
    def prepare(self, sd: Dict[str, torch.Tensor]) -> None:
        """sd: ImageEncoderViT.state_dict() (keys without the ``image_encoder.`` prefix), tensors on self.device."""
        D, dev = self.D, self.device
        w = self.w
        g = lambda k: sd[k].to(dev)
        w["patch_w"] = _bf(g("patch_embed.proj.weight").reshape(D, 768))
        w["patch_b"] = _f32(g("patch_embed.proj.bias"))
        w["pos"] = _f32(g("pos_embed").reshape(NTOK, D))
        w["hfc_w"] = _bf(g("hfc_embed.proj.weight").reshape(HFC, 256))
        w["hfc_b"] = _f32(g("hfc_embed.proj.bias"))
        a = "hfc_attn."
        w["ph_w"] = _bf(g(a + "proj_hfc.weight").reshape(HFC, HFC))
        w["ph_b"] = _f32(g(a + "proj_hfc.bias"))
        w["pos_hfc"] = _f32(g(a + "pos_embed").reshape(HFC, NTOK).t())  # NCHW -> [token, channel]
        w["pp_w"] = _bf(g(a + "proj_patch.weight").reshape(HFC, D))
        w["pp_b"] = _f32(g(a + "proj_patch.bias"))
        wi, bi = g(a + "cross_attn.in_proj_weight"), g(a + "cross_attn.in_proj_bias")
        w["q_w"], w["q_b"] = _bf(wi[:HFC]), _f32(bi[:HFC])
        w["kv_w"], w["kv_b"] = _bf(wi[HFC:]), _f32(bi[HFC:])
        w["o_w"], w["o_b"] = _bf(g(a + "cross_attn.out_proj.weight")), _f32(g(a + "cross_attn.out_proj.bias"))
        for n in ("linear1", "linear2"):
            w[n + "_w"], w[n + "_b"] = _bf(g(a + n + ".weight")), _f32(g(a + n + ".bias"))
        for n in ("norm1", "norm2"):
            w["hfc_" + n + "_g"], w["hfc_" + n + "_b"] = _f32(g(a + n + ".weight")), _f32(g(a + n + ".bias"))
        w["back_w"] = _bf(g(a + "proj_back.weight").reshape(D, HFC))
        w["back_b"] = _f32(g(a + "proj_back.bias"))
        # low-pass DFT operator (constant): GEMM1 weight rows interleaved (x', part), GEMM2 weight [y', (part, y)]
        Lr, Li = lowpass_operator_tables()
        lp1 = torch.stack([Lr, Li], dim=1).reshape(2048, 1024).to(dev)
        lp2 = torch.cat([Lr, -Li], dim=1).to(dev)
        w["lp1"], w["lp2"] = _bf(lp1), _bf(lp2)
        if self.hfc_precision == "split":  # [Wh | Wh | Wl] against the operand rows [hi | lo | hi]
            for n, full in (("lp1", lp1), ("lp2", lp2)):
                hi = w[n]
                lo = _bf(full - hi.to(torch.float64))
                w[n + "s"] = torch.cat([hi, hi, lo], dim=1).contiguous()
        for i in range(self.depth):
            b = f"blocks.{i}."
            p = f"b{i}."
            qkv_b = _f32(g(b + "attn.qkv.bias")).clone()
            b_v = qkv_b[2 * D:].clone()
            # k bias cancels in softmax, v bias moves to the output (sum p = 1): see csrc/attn_window.cu
            qkv_b[D:] = 0
            w[p + "qkv_w"], w[p + "qkv_b"] = _bf(g(b + "attn.qkv.weight")), qkv_b
            pw = g(b + "attn.proj.weight").float()
            w[p + "proj_w"] = _bf(pw)
            w[p + "proj_b"] = _f32(g(b + "attn.proj.bias").float() + pw @ b_v)
            rh, rw = g(b + "attn.rel_pos_h"), g(b + "attn.rel_pos_w")
            if i in self.glob:
                t = torch.zeros(256, self.hd, device=dev, dtype=torch.bfloat16)
                t[:127], t[128:255] = rh.to(torch.bfloat16), rw.to(torch.bfloat16)
            else:
                t = torch.zeros(64, self.hd, device=dev, dtype=torch.bfloat16)
                t[:27], t[32:59] = rh.to(torch.bfloat16), rw.to(torch.bfloat16)
            w[p + "rel"] = t
            for n in ("norm1", "norm2"):
                w[p + n + "_g"], w[p + n + "_b"] = _f32(g(b + n + ".weight")), _f32(g(b + n + ".bias"))
            w[p + "lin1_w"], w[p + "lin1_b"] = _bf(g(b + "mlp.lin1.weight")), _f32(g(b + "mlp.lin1.bias"))
            w[p + "lin2_w"], w[p + "lin2_b"] = _bf(g(b + "mlp.lin2.weight")), _f32(g(b + "mlp.lin2.bias"))
        w["neck0_w"] = _bf(g("neck.0.weight").reshape(256, D))
        w["neck1_g"], w["neck1_b"] = _f32(g("neck.1.weight")), _f32(g("neck.1.bias"))
        w["neck2_w"] = _bf(g("neck.2.weight").permute(0, 2, 3, 1).reshape(256, 9 * 256))
        w["neck3_g"], w["neck3_b"] = _f32(g("neck.3.weight")), _f32(g("neck.3.bias"))

    # ------------------------------------------------------------------ stages
    def highpass(self, img: torch.Tensor, want_image: bool = False):
        """E0 (+ im2col for E1/E2). Returns (patch rows bf16 [M,768], hfc rows bf16 [M,256], hfc image or None)."""
        B = img.shape[0]
        ws, w = self.ws, self.w
        self.gen_rows += 1
        a_patch = ws.get("a_patch", (B * NTOK, 768), torch.bfloat16)
        low_t = ws.get("lp_low", (B * 1024, 1024), torch.float32)
        if self.hfc_precision == "split":
            gray = ws.get("gray", (B * 1024, 3072), torch.bfloat16)
            ops.patchify(img, a_patch, gray, True)
            p1 = ws.get("lp_p1", (B * 1024, 2048), torch.float32)
            p1t = ws.get("lp_p1t", (B * 1024, 6144), torch.bfloat16)
            with credit(1.0 / 3.0):  # three executed products per algorithmic one
                _gemm(gray, w["lp1s"], out_f32=p1)
                ops.transpose_split(p1.view(B, 1024, 2048), p1t)
                _gemm(p1t, w["lp2s"], out_f32=low_t)
        else:
            gray = ws.get("gray", (B * 1024, 1024), torch.bfloat16)
            ops.patchify(img, a_patch, gray)
            p1 = ws.get("lp_p1", (B * 1024, 2048), torch.bfloat16)
            _gemm(gray, w["lp1"], out_bf16=p1)
            p1t = ws.get("lp_p1t", (B, 2048, 1024), torch.bfloat16)
            ops.transpose(p1.view(B, 1024, 2048), p1t)
            _gemm(p1t.view(B * 1024, 2048), w["lp2"], out_f32=low_t)
        a_hfc = ws.get("a_hfc", (B * NTOK, 256), torch.bfloat16)
        hfc_img = torch.empty(B, 1, 1024, 1024, device=img.device, dtype=torch.float32) if want_image else None
        ops.hfc_finalize(img, low_t, a_hfc, hfc_img)
        self.launches += 5
        return a_patch, a_hfc, hfc_img

    def patch_rows(self, img: torch.Tensor, hfc_img: torch.Tensor):
        """im2col rows of a tile batch and of a caller-supplied high-pass image (ImageEncoderViT.forward called with an
        x_hfc that did not come from MedSAM.fft in the same pass)."""
        B = img.shape[0]
        self.gen_rows += 1
        a_patch = self.ws.get("a_patch", (B * NTOK, 768), torch.bfloat16)
        ops.patchify(img, a_patch, None)
        a_hfc = self.ws.get("a_hfc", (B * NTOK, 256), torch.bfloat16)
        ops.patchify(hfc_img, a_hfc, None)
        self.launches += 2
        return a_patch, a_hfc

    def encode(self, a_patch: torch.Tensor, a_hfc: torch.Tensor, B: int, taps: Optional[dict] = None):
        """E1..E12 on im2col rows. Returns NHWC features: (fp32 [M,256], bf16 [M,256])."""
        D, H, M = self.D, self.H, B * NTOK
        self.gen_feat += 1
        ws, w = self.ws, self.w
        bf, f32 = torch.bfloat16, torch.float32
        n0 = self.launches
        # E1: patch embed + pos embed  -> residual stream (fp32) and its bf16 copy
        x = ws.get("x", (M, D), f32)
        xb = ws.get("xb", (M, D), bf)
        _gemm(a_patch, w["patch_w"], w["patch_b"], w["pos"], NTOK, xb, x)
        # E2: hfc embed, E3: HFC cross-attention branch
        h0 = ws.get("h0", (M, HFC), bf)
        _gemm(a_hfc, w["hfc_w"], w["hfc_b"], out_bf16=h0)
        hk = ws.get("hk", (M, HFC), bf)
        _gemm(h0, w["ph_w"], w["ph_b"], w["pos_hfc"], NTOK, hk)
        pq = ws.get("pq", (M, HFC), f32)
        pqb = ws.get("pqb", (M, HFC), bf)
        _gemm(xb, w["pp_w"], w["pp_b"], out_bf16=pqb, out_f32=pq)
        qh = ws.get("h0", (M, HFC), bf)  # h0 is dead after proj_hfc
        _gemm(pqb, w["q_w"], w["q_b"], out_bf16=qh)
        kvh = ws.get("kvh", (M, 2 * HFC), bf)
        _gemm(hk, w["kv_w"], w["kv_b"], out_bf16=kvh)
        ao = ws.get("hk", (M, HFC), bf)  # hk is dead after the kv projection
        ops.attn_flash(qh, 0, kvh, 0, kvh, HFC, None, ao, B, 8, NTOK, NTOK, 128, 1.0 / math.sqrt(128.0))
        y1 = ws.get("y1", (M, HFC), f32)
        _gemm(ao, w["o_w"], w["o_b"], pq, M, out_f32=y1)
        z1 = ws.get("pq", (M, HFC), f32)
        z1b = ws.get("pqb", (M, HFC), bf)
        ops.layernorm(y1, w["hfc_norm1_g"], w["hfc_norm1_b"], z1b, z1, None, 0, None, 1e-5)
        f1 = ws.get("h0", (M, HFC), bf)
        _gemm(z1b, w["linear1_w"], w["linear1_b"], out_bf16=f1, act=ACT_RELU)
        _gemm(f1, w["linear2_w"], w["linear2_b"], z1, M, out_f32=y1)
        z2b = ws.get("pqb", (M, HFC), bf)
        ops.layernorm(y1, w["hfc_norm2_g"], w["hfc_norm2_b"], z2b, None, None, 0, None, 1e-5)
        # image_encoder.py:512: RAW reinterpret [N,1024] -> [1024,N] per image: the GEMM A
        # operand is the transpose of that view, so the 1x1 proj_back contracts over token groups.
        z2t = ws.get("h0", (B, NTOK, HFC), bf)
        ops.transpose(z2b.view(B, HFC, NTOK), z2t)
        _gemm(z2t.view(M, HFC), w["back_w"], w["back_b"], x, M, out_f32=x)  # E4: x = x_hfc + x (in place)
        self.launches += 14
        if taps is not None:
            taps["after_hfc"] = x.clone()
        # E5..E11: blocks
        xn = ws.get("xb", (M, D), bf)
        qkv = ws.get("qkv", (M, 3 * D), bf)
        att = ws.get("att", (M, D), bf)
        hid = ws.get("hid", (M, 4 * D), bf)
        scale = self.hd ** -0.5
        for i in range(self.depth):
            p = f"b{i}."
            ops.layernorm(x, w[p + "norm1_g"], w[p + "norm1_b"], xn, None, None, 0, None, 1e-6)
            _gemm(xn, w[p + "qkv_w"], w[p + "qkv_b"], out_bf16=qkv)
            if i in self.glob:
                ops.attn_flash(qkv, 0, qkv, D, qkv, 2 * D, w[p + "rel"], att, B, H, NTOK, NTOK, self.hd, scale)
            else:
                ops.attn_window(qkv, w[p + "rel"], att, H, scale)
            _gemm(att, w[p + "proj_w"], w[p + "proj_b"], x, M, out_f32=x)
            ops.layernorm(x, w[p + "norm2_g"], w[p + "norm2_b"], xn, None, None, 0, None, 1e-6)
            _gemm(xn, w[p + "lin1_w"], w[p + "lin1_b"], out_bf16=hid, act=ACT_GELU)
            last = i == self.depth - 1
            _gemm(hid, w[p + "lin2_w"], w[p + "lin2_b"], x, M, out_bf16=xn if last else None, out_f32=x)
            self.launches += 7
            if taps is not None:
                taps[f"block{i}"] = x.clone()
        # E12: neck (NHWC): 1x1 -> LN2d -> 3x3 -> LN2d
        t0 = ws.get("t0", (M, 256), f32)
        _gemm(xn, w["neck0_w"], out_f32=t0)
        t1 = ws.get("t1", (M, 256), bf)
        ops.layernorm(t0, w["neck1_g"], w["neck1_b"], t1, None, None, 0, None, 1e-6)
        ops.conv3x3(t1.view(B, GRID, GRID, 256), w["neck2_w"], None, t0)
        feat = ws.get("feat", (M, 256), f32)
        featb = ws.get("featb", (M, 256), bf)
        ops.layernorm(t0, w["neck3_g"], w["neck3_b"], featb, feat, None, 0, None, 1e-6)
        self.launches += 4
        del n0
        return feat, featb

    def to_nchw(self, feat: torch.Tensor, B: int) -> torch.Tensor:
        out = torch.empty(B, 256, GRID, GRID, device=feat.device, dtype=torch.float32)
        ops.transpose(feat.view(B, NTOK, 256), out.view(B, 256, NTOK))
        self.launches += 1
        return out


class DecoderEngine:
    """Derived weights + schedule for MaskDecoder / TwoWayTransformer / heads."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.ws = _Workspace(self.device)
        self.w: Dict[str, torch.Tensor] = {}
        self.launches = 0

    @staticmethod
    def dense_pe_tokens(gaussian_matrix: torch.Tensor) -> torch.Tensor:
        """D1: PositionEmbeddingRandom over the 64x64 grid (pos_encoder.py:24-33,50-70), token-major [4096,256].
        Constant per model: evaluated once at load time, not part of the per-tile path."""
        G = gaussian_matrix.float()
        c = (torch.arange(GRID, device=G.device, dtype=torch.float32) + 0.5) / GRID
        xy = torch.stack([c[None, :].expand(GRID, GRID), c[:, None].expand(GRID, GRID)], dim=-1)
        ang = 2 * math.pi * ((2 * xy - 1) @ G)
        return torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1).reshape(NTOK, 256).contiguous()

    def prepare_transformer(self, sd: Dict[str, torch.Tensor], prefix: str = "") -> None:
        """sd: TwoWayTransformer.state_dict() (prefix '') or MaskDecoder.state_dict() (prefix 'transformer.')."""
        dev, w = self.device, self.w
        g = lambda k: sd[prefix + k].to(dev)

        def attn(dst: str, src: str) -> None:
            for n in ("q", "k", "v", "out"):
                w[f"{dst}.{n}_w"] = _bf(g(f"{src}.{n}_proj.weight"))
                w[f"{dst}.{n}_b"] = _f32(g(f"{src}.{n}_proj.bias"))
            # head-padded copies for the tcgen05 flash kernel (large query counts): every head of 16 / 32 channels gets its
            # own 64-column slot, the padding rows / columns are zero, so q.k and the out projection are unchanged
            C = w[f"{dst}.q_w"].shape[0]
            hd = C // 8
            idx = (torch.arange(8, device=dev)[:, None] * 64 + torch.arange(hd, device=dev)[None, :]).reshape(-1)
            for n in ("q", "k", "v"):
                wp = torch.zeros(512, 256, device=dev, dtype=torch.bfloat16)
                bp = torch.zeros(512, device=dev, dtype=torch.float32)
                wp[idx], bp[idx] = w[f"{dst}.{n}_w"], w[f"{dst}.{n}_b"]
                w[f"{dst}.{n}p_w"], w[f"{dst}.{n}p_b"] = wp, bp
            wo = torch.zeros(256, 512, device=dev, dtype=torch.bfloat16)
            wo[:, idx] = w[f"{dst}.out_w"]
            w[f"{dst}.outp_w"] = wo

        for i in range(2):
            L = f"layers.{i}."
            attn(f"l{i}.self", L + "self_attn")
            attn(f"l{i}.t2i", L + "cross_attn_token_to_image")
            attn(f"l{i}.i2t", L + "cross_attn_image_to_token")
            for n in ("norm1", "norm2", "norm3", "norm4"):
                w[f"l{i}.{n}_g"], w[f"l{i}.{n}_b"] = _f32(g(L + n + ".weight")), _f32(g(L + n + ".bias"))
            w[f"l{i}.lin1_w"], w[f"l{i}.lin1_b"] = _bf(g(L + "mlp.lin1.weight")), _f32(g(L + "mlp.lin1.bias"))
            w[f"l{i}.lin2_w"], w[f"l{i}.lin2_b"] = _bf(g(L + "mlp.lin2.weight")), _f32(g(L + "mlp.lin2.bias"))
        attn("final", "final_attn_token_to_image")
        w["nf_g"], w["nf_b"] = _f32(g("norm_final_attn.weight")), _f32(g("norm_final_attn.bias"))

    def prepare_heads(self, sd: Dict[str, torch.Tensor]) -> None:
        """sd: MaskDecoder.state_dict()."""
        dev, w = self.device, self.w
        w["tokens"] = _f32(sd["mask_tokens.weight"].to(dev))
        for head in ("class_embed", "bbox_embed"):
            for i in range(3):
                w[f"{head}.{i}_w"] = _bf(sd[f"{head}.layers.{i}.weight"].to(dev))
                w[f"{head}.{i}_b"] = _f32(sd[f"{head}.layers.{i}.bias"].to(dev))

    FLASH_MIN_TOKENS = 256  # query counts from here on go through the tcgen05 flash kernel (head dims padded to 64)

    def _attention(self, pre: str, q_in, k_in, v_in, B: int, Tq: int, Tk: int, name: str):
        """transformer.py:218-240: projections -> 8-head attention.  Returns (pre-out_proj rows bf16, out_proj weight).
        Small query counts (the reference's 51): thread-per-query kernel on the 16 / 32-channel heads.  Large ones (dense
        herds, 900 queries): the projections write head-padded [*, 8 x 64] rows and the v4 flash kernel does the rest."""
        w, ws = self.w, self.ws
        bf = torch.bfloat16
        C = w[pre + ".q_w"].shape[0]
        hd = C // 8
        flash = min(Tq, Tk) >= self.FLASH_MIN_TOKENS
        sfx, Cp = ("p", 512) if flash else ("", C)
        qh = ws.get(name + ".qh" + sfx, (B * Tq, Cp), bf)
        kh = ws.get(name + ".kh" + sfx, (B * Tk, Cp), bf)
        vh = ws.get(name + ".vh" + sfx, (B * Tk, Cp), bf)
        _gemm(q_in, w[f"{pre}.q{sfx}_w"], w[f"{pre}.q{sfx}_b"], out_bf16=qh)
        _gemm(k_in, w[f"{pre}.k{sfx}_w"], w[f"{pre}.k{sfx}_b"], out_bf16=kh)
        _gemm(v_in, w[f"{pre}.v{sfx}_w"], w[f"{pre}.v{sfx}_b"], out_bf16=vh)
        o = ws.get(name + ".o" + sfx, (B * Tq, Cp), bf)
        if flash:
            ops.attn_flash(qh, 0, kh, 0, vh, 0, None, o, B, 8, Tq, Tk, 64, 1.0 / math.sqrt(hd))
        else:
            ops.attn_small(qh, kh, vh, o, B, 8, Tq, Tk, hd, 1.0 / math.sqrt(hd))
        self.launches += 4
        return o, w[f"{pre}.out{sfx}_w"]

    def transformer(self, feat: torch.Tensor, pe: torch.Tensor, tokens: torch.Tensor, B: int, Q: int):
        """TwoWayTransformer.forward (transformer.py:62-106) on token-major rows.

        feat fp32 [B*4096,256] (image embedding), pe fp32 [4096,256] (shared) or [B*4096,256], tokens fp32 [Q,256]
        (batch-broadcast point embedding) or [B*Q,256].  Returns (queries fp32 [B*Q,256], its bf16 copy, keys fp32).
        """
        w, ws = self.w, self.ws
        bf, f32 = torch.bfloat16, torch.float32
        MQ, MK = B * Q, B * NTOK
        tmod, pmod = tokens.shape[0], pe.shape[0]
        X = ws.get("X", (MQ, 256), f32)
        Xb = ws.get("Xb", (MQ, 256), bf)
        Xpe = ws.get("Xpe", (MQ, 256), bf)
        keys = ws.get("keys", (MK, 256), f32)
        keysb = ws.get("keysb", (MK, 256), bf)
        keyspe = ws.get("keyspe", (MK, 256), bf)
        Y = ws.get("Y", (MQ, 256), f32)
        ops.add_cast(None, tokens, tmod, Xb)  # queries = point_embedding (bf16 operand copy, batch-broadcast)
        ops.add_cast(feat, None, 0, keysb)
        ops.add_cast(feat, pe, pmod, keyspe)
        keys_f32 = feat
        self.launches += 4
        for i in range(2):
            L = f"l{i}"
            if i == 0:  # skip_first_layer_pe: q = k = v = queries, output REPLACES the queries (transformer.py:155-156)
                o, ow = self._attention(L + ".self", Xb, Xb, Xb, B, Q, Q, "self")
                _gemm(o, ow, w[L + ".self.out_b"], out_f32=Y)
            else:
                o, ow = self._attention(L + ".self", Xpe, Xpe, Xb, B, Q, Q, "self")
                _gemm(o, ow, w[L + ".self.out_b"], X, MQ, out_f32=Y)
            ops.layernorm(Y, w[L + ".norm1_g"], w[L + ".norm1_b"], Xb, X, tokens, tmod, Xpe, 1e-5)
            # tokens -> image cross attention
            o, ow = self._attention(L + ".t2i", Xpe, keyspe, keysb, B, Q, NTOK, "t2i")
            _gemm(o, ow, w[L + ".t2i.out_b"], X, MQ, out_f32=Y)
            ops.layernorm(Y, w[L + ".norm2_g"], w[L + ".norm2_b"], Xb, X, None, 0, None, 1e-5)
            # MLP
            hid = ws.get("hid", (MQ, 2048), bf)
            _gemm(Xb, w[L + ".lin1_w"], w[L + ".lin1_b"], out_bf16=hid, act=ACT_RELU)
            _gemm(hid, w[L + ".lin2_w"], w[L + ".lin2_b"], X, MQ, out_f32=Y)
            ops.layernorm(Y, w[L + ".norm3_g"], w[L + ".norm3_b"], Xb, X, tokens, tmod, Xpe, 1e-5)
            # image -> tokens cross attention (updates the keys)
            o, ow = self._attention(L + ".i2t", keyspe, Xpe, Xb, B, NTOK, Q, "i2t")
            ky = ws.get("keysY", (MK, 256), f32)
            _gemm(o, ow, w[L + ".i2t.out_b"], keys_f32, MK, out_f32=ky)
            ops.layernorm(ky, w[L + ".norm4_g"], w[L + ".norm4_b"], keysb, keys, pe, pmod, keyspe, 1e-5)
            keys_f32 = keys
            self.launches += 9
        o, ow = self._attention("final", Xpe, keyspe, keysb, B, Q, NTOK, "t2i")
        _gemm(o, ow, w["final.out_b"], X, MQ, out_f32=Y)
        hs = ws.get("hs", (MQ, 256), f32)
        hsb = ws.get("hsb", (MQ, 256), bf)
        ops.layernorm(Y, w["nf_g"], w["nf_b"], hsb, hs, None, 0, None, 1e-5)
        self.launches += 2
        return hs, hsb, keys_f32

    def heads(self, hsb: torch.Tensor, B: int, Q: int):
        """class_embed / bbox_embed MLPs (box_decoder.py:102-103,154-176). Returns fp32 logits [B,Q,8], boxes [B,Q,4]."""
        w, ws = self.w, self.ws
        MQ = B * Q
        logits = torch.empty(B, Q, 8, device=hsb.device, dtype=torch.float32)
        boxes = torch.empty(B, Q, 4, device=hsb.device, dtype=torch.float32)
        h1 = ws.get("head1", (MQ, 256), torch.bfloat16)
        h2 = ws.get("head2", (MQ, 256), torch.bfloat16)
        for head, out, act in (("class_embed", logits, ACT_NONE), ("bbox_embed", boxes, ACT_SIGMOID)):
            _gemm(hsb, w[head + ".0_w"], w[head + ".0_b"], out_bf16=h1, act=ACT_RELU)
            _gemm(h1, w[head + ".1_w"], w[head + ".1_b"], out_bf16=h2, act=ACT_RELU)
            _gemm(h2, w[head + ".2_w"], w[head + ".2_b"], out_f32=out.view(MQ, -1), act=act)
        self.launches += 6
        return logits, boxes
