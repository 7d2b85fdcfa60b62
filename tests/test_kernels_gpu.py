"""Kernel-level parity tests on the B200: every C-ABI entry point against a plain fp32 torch / numpy restatement
of the same op on the same seeded inputs.  Integer outputs must be bit-exact; float tolerances are stated inline."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # collected on CPU boxes too; -m gpu selects them only on the GPU box
    pytest.skip("needs a GPU", allow_module_level=True)

from wildlifemapper_b200.ops import ops  # noqa: E402
from oracle import post as opost  # noqa: E402

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def rel_err(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-6)).item()


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 0), (256, 256, 128, 0), (1000, 768, 768, 0), (4096, 2304, 768, 0), (4096, 768, 3072, 0),
    (384, 128, 256, 0), (51, 8, 256, 0), (300, 4, 256, 0), (1632, 2048, 256, 0), (1632, 256, 2048, 0),
    (512, 1024, 1024, 128), (512, 1280, 1280, 0), (640, 192, 128, 64), (20000, 256, 768, 0),
    # bn = 512: the CTA-pair kernel (256 x 256 tile per cluster of two CTAs), incl. ragged M and more tiles than clusters
    (256, 256, 128, 512), (1000, 768, 768, 512), (4096, 2304, 768, 512), (300, 512, 64, 512), (20000, 256, 768, 512),
    (40000, 1024, 256, 512), (32768, 768, 3072, 0),
])
def test_gemm_matches_fp32(M, N, K, bn):
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, w, None, None, 0, None, out, 0, bn)
    ref = a.float() @ w.float().t()
    assert rel_err(out, ref) < 2e-3  # bf16 products are exact in fp32; only accumulation order differs


@pytest.mark.parametrize("bn", [0, 512])
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm_epilogues(act, bn):
    M, N, K = 700, 768, 256
    a, w = rnd(M, K, seed=3), rnd(N, K, seed=4, scale=K ** -0.5)
    bias = rnd(N, seed=5, dtype=torch.float32)
    res = rnd(100, N, seed=6, dtype=torch.float32)
    o32 = torch.empty(M, N, device=DEV, dtype=torch.float32)
    o16 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, bias, res, 100, o16, o32, act, bn)
    y = a.float() @ w.float().t() + bias
    y = [y, torch.nn.functional.gelu(y), torch.relu(y), torch.sigmoid(y)][act]
    y = y + res[torch.arange(M, device=DEV) % 100]
    assert (o32 - y).abs().max().item() < 5e-3
    assert (o16.float() - y).abs().max().item() < 5e-2  # bf16 rounding of values up to ~6


def test_gemm_strided_views():
    # A and the outputs are column slices of wider buffers (how qkv / kv projections are consumed)
    M, N, K = 512, 256, 128
    big = rnd(M, 3 * K, seed=7)
    a = big[:, K:2 * K]
    w = rnd(N, K, seed=8, scale=K ** -0.5)
    outbig = torch.zeros(M, 2 * N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, None, None, 0, outbig[:, N:], None, 0, 0)
    ref = a.float() @ w.float().t()
    assert rel_err(outbig[:, N:], ref) < 1e-2
    assert outbig[:, :N].abs().max().item() == 0


def test_conv3x3_matches_conv2d():
    B, C, N = 2, 256, 256
    x = rnd(B, 64, 64, C, seed=9)
    wt = rnd(N, C, 3, 3, seed=10, scale=(9 * C) ** -0.5)
    w2 = wt.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()
    out = torch.empty(B * 4096, N, device=DEV, dtype=torch.float32)
    ops.conv3x3(x, w2, None, out)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
    assert rel_err(out.view(B, 64, 64, N), ref) < 2e-3


# ------------------------------------------------------------------ bandwidth kernels
@pytest.mark.parametrize("D", [128, 256, 640, 768, 1024, 1280])
def test_layernorm(D):
    rows = 1000
    x = rnd(rows, D, seed=11, scale=3.0, dtype=torch.float32) + 0.5
    g, b = rnd(D, seed=12, dtype=torch.float32), rnd(D, seed=13, dtype=torch.float32)
    add = rnd(64, D, seed=14, dtype=torch.float32)
    y16 = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
    y32 = torch.empty(rows, D, device=DEV, dtype=torch.float32)
    y2 = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
    ops.layernorm(x, g, b, y16, y32, add, 64, y2, 1e-6)
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-6)
    assert (y32 - ref).abs().max().item() < 2e-5
    assert (y16.float() - ref).abs().max().item() <= ref.abs().max().item() * 2 ** -8
    ref2 = ref + add[torch.arange(rows, device=DEV) % 64]
    assert (y2.float() - ref2).abs().max().item() <= ref2.abs().max().item() * 2 ** -8


def test_patchify_and_gray():
    B = 2
    img = rnd(B, 3, 1024, 1024, seed=15, dtype=torch.float32)
    patches = torch.empty(B * 4096, 768, device=DEV, dtype=torch.bfloat16)
    gray = torch.empty(B, 1024, 1024, device=DEV, dtype=torch.bfloat16)
    ops.patchify(img, patches, gray)
    ref = torch.nn.functional.unfold(img, 16, stride=16).transpose(1, 2).reshape(B * 4096, 768)  # (c,ky,kx) order
    assert torch.equal(patches, ref.to(torch.bfloat16))
    g = 0.2989 * img[:, 0] + 0.587 * img[:, 1] + 0.114 * img[:, 2]
    assert (gray.float() - g).abs().max().item() < 0.02


def test_patchify_gray_split():
    """gray_split: rows [hi | lo | hi] with hi = bf16(g) (bit-identical to the unsplit plane), hi + lo = g to ~2^-17."""
    B = 2
    img = rnd(B, 3, 1024, 1024, seed=15, dtype=torch.float32)
    patches = torch.empty(B * 4096, 768, device=DEV, dtype=torch.bfloat16)
    gray = torch.empty(B, 1024, 1024, device=DEV, dtype=torch.bfloat16)
    ops.patchify(img, patches, gray)
    patches2 = torch.empty_like(patches)
    gs = torch.empty(B * 1024, 3072, device=DEV, dtype=torch.bfloat16)
    ops.patchify(img, patches2, gs, True)
    assert torch.equal(patches, patches2)
    assert torch.equal(gs[:, :1024], gray.view(B * 1024, 1024)) and torch.equal(gs[:, 2048:], gs[:, :1024])
    g = (0.2989 * img[:, 0] + 0.587 * img[:, 1] + 0.114 * img[:, 2]).view(B * 1024, 1024)
    err = (gs[:, :1024].float() + gs[:, 1024:2048].float() - g).abs().max().item()
    assert err <= g.abs().max().item() * 2 ** -15, err


def test_transpose_split():
    """fp32 [b, R, C] -> out[b][c / 2][seg][c % 2][r] with seg (hi, lo, hi): bit-exact against the same split in torch."""
    b, R, C = 2, 128, 192
    x = rnd(b, R, C, seed=17, dtype=torch.float32) * 3
    out = torch.empty(b * (C // 2), 6 * R, device=DEV, dtype=torch.bfloat16)
    ops.transpose_split(x, out)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)

    def lay(t):  # [b, R, C] -> [b, C/2, 2, R]
        return t.transpose(1, 2).reshape(b, C // 2, 2, R)
    ref = torch.stack([lay(hi), lay(lo), lay(hi)], dim=2).reshape(b * (C // 2), 6 * R)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("R,C,dt", [(1024, 4096, torch.bfloat16), (4096, 256, torch.float32), (100, 70, torch.float32),
                                     (1024, 2048, torch.bfloat16), (100, 70, torch.bfloat16), (67, 33, torch.bfloat16)])
def test_transpose(R, C, dt):
    x = rnd(3, R, C, seed=16, dtype=dt)
    out = torch.empty(3, C, R, device=DEV, dtype=dt)
    ops.transpose(x, out)
    assert torch.equal(out, x.transpose(1, 2).contiguous())


def test_hfc_finalize():
    B = 2
    img = rnd(B, 3, 1024, 1024, seed=17, dtype=torch.float32)
    low = rnd(B, 1024, 1024, seed=18, dtype=torch.float32)  # low[b][y][x]
    low_t = low.transpose(1, 2).contiguous()
    patches = torch.empty(B * 4096, 256, device=DEV, dtype=torch.bfloat16)
    himg = torch.empty(B, 1024, 1024, device=DEV, dtype=torch.float32)
    ops.hfc_finalize(img, low_t, patches, himg)
    g = 0.2989 * img[:, 0] + 0.587 * img[:, 1] + 0.114 * img[:, 2]
    ref = (g - low).abs()
    assert (himg - ref).abs().max().item() < 1e-5
    refp = torch.nn.functional.unfold(ref[:, None], 16, stride=16).transpose(1, 2).reshape(B * 4096, 256)
    assert (patches.float() - refp).abs().max().item() <= refp.abs().max().item() * 2 ** -8


def test_add_cast():
    a = rnd(500, 256, seed=19, dtype=torch.float32)
    b = rnd(50, 256, seed=20, dtype=torch.float32)
    out = torch.empty(500, 256, device=DEV, dtype=torch.bfloat16)
    ops.add_cast(a, b, 50, out)
    assert torch.equal(out, (a + b[torch.arange(500, device=DEV) % 50]).to(torch.bfloat16))
    ops.add_cast(a, None, 0, out)
    assert torch.equal(out, a.to(torch.bfloat16))


# ------------------------------------------------------------------ attention
def ref_attention(q, k, v, scale, bias=None):
    s = (q.float() * scale) @ k.float().transpose(-1, -2)
    if bias is not None:
        s = s + bias
    return torch.softmax(s, -1) @ v.float()


@pytest.mark.parametrize("Tq,Tk,hd,H", [(51, 51, 32, 8), (51, 4096, 16, 8), (4096, 51, 16, 8), (900, 900, 32, 8), (7, 130, 16, 2)])
def test_attn_small(Tq, Tk, hd, H):
    B = 2
    q, k, v = rnd(B * Tq, H * hd, seed=21), rnd(B * Tk, H * hd, seed=22), rnd(B * Tk, H * hd, seed=23)
    out = torch.empty(B * Tq, H * hd, device=DEV, dtype=torch.bfloat16)
    scale = 1 / math.sqrt(hd)
    ops.attn_small(q, k, v, out, B, H, Tq, Tk, hd, scale)
    sp = lambda t, T: t.view(B, T, H, hd).transpose(1, 2)
    ref = ref_attention(sp(q, Tq), sp(k, Tk), sp(v, Tk), scale).transpose(1, 2).reshape(B * Tq, H * hd)
    assert (out.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("hd,H,T", [(64, 2, 256), (64, 3, 4096), (128, 2, 512), (128, 8, 4096), (80, 3, 1024)])
def test_attn_flash_plain(hd, H, T, flash_version):
    B = 2 if T <= 512 else 1
    q = rnd(B * T, H * hd, seed=24)
    kv = rnd(B * T, 2 * H * hd, seed=25)
    out = torch.zeros(B * T, H * hd, device=DEV, dtype=torch.bfloat16)
    scale = 1 / math.sqrt(hd)
    ops.attn_flash(q, 0, kv, 0, kv, H * hd, None, out, B, H, T, T, hd, scale)
    sp = lambda t: t.reshape(B, T, H, hd).transpose(1, 2)
    ref = ref_attention(sp(q), sp(kv[:, :H * hd]), sp(kv[:, H * hd:]), scale).transpose(1, 2).reshape(B * T, H * hd)
    # P is rounded to bf16 before the P.V MMA and the output to bf16: ~2^-8 relative
    assert (out.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("Tq,Tk,hd", [(900, 900, 64), (900, 4096, 64), (4096, 900, 64), (51, 130, 64), (300, 77, 128)])
def test_attn_flash_ragged(Tq, Tk, hd):
    """Query / key counts that are not multiples of the tile sizes (decoder attention with 900 queries): rows beyond Tq
    are not stored, keys beyond Tk are masked (v4 kernel)."""
    B, H = 2, 8
    q = rnd(B * Tq, H * hd, seed=50)
    k, v = rnd(B * Tk, H * hd, seed=51), rnd(B * Tk, H * hd, seed=52)
    out = torch.full((B * Tq + 7, H * hd), 7.0, device=DEV, dtype=torch.bfloat16)  # canary rows behind the last query
    scale = 0.25
    ops.attn_flash(q, 0, k, 0, v, 0, None, out, B, H, Tq, Tk, hd, scale)
    sp = lambda t, T: t.view(B, T, H, hd).transpose(1, 2)
    ref = ref_attention(sp(q, Tq), sp(k, Tk), sp(v, Tk), scale).transpose(1, 2).reshape(B * Tq, H * hd)
    assert (out[:B * Tq].float() - ref).abs().max().item() < 2e-2
    assert (out[B * Tq:] == 7.0).all()


@pytest.mark.parametrize("hd", [64, 128])
def test_attn_flash_rising_maxima(hd, flash_version):
    """Scores that keep growing along the key axis force the lazy-rescale path (reference maximum raised, O and l
    rescaled in TMEM) many times per row, including late key tiles; the two query tiles of a CTA see different
    score ranges, so the two softmax warpgroups drift apart in time."""
    B, H, T = 1, 2, 1024
    q = rnd(B * T, H * hd, seed=40, scale=2.0)
    kv = rnd(B * T, 2 * H * hd, seed=41)
    ramp = torch.linspace(0.2, 4.0, T, device=DEV)[:, None]
    kv[:, :H * hd] = (kv[:, :H * hd].float() * ramp).to(torch.bfloat16)
    q[T // 4: T // 2] = (q[T // 4: T // 2].float() * 0.05).to(torch.bfloat16)  # a flat-score query tile next to a peaky one
    out = torch.zeros(B * T, H * hd, device=DEV, dtype=torch.bfloat16)
    scale = 1 / math.sqrt(hd)
    ops.attn_flash(q, 0, kv, 0, kv, H * hd, None, out, B, H, T, T, hd, scale)
    sp = lambda t: t.reshape(B, T, H, hd).transpose(1, 2)
    ref = ref_attention(sp(q), sp(kv[:, :H * hd]), sp(kv[:, H * hd:]), scale).transpose(1, 2).reshape(B * T, H * hd)
    assert (out.float() - ref).abs().max().item() < 3e-2


FLASH_VERSIONS = [7, 4]  # selectable flash-attention kernel generations (wm_set_flash_version); the first one is the default
FLASH_DEFAULT = FLASH_VERSIONS[0]


@pytest.fixture(params=FLASH_VERSIONS)
def flash_version(request):
    """Every selectable flash-attention kernel generation stays parity-checked."""
    from wildlifemapper_b200 import lib
    lib.call("wm_set_flash_version", request.param)
    yield request.param
    lib.call("wm_set_flash_version", FLASH_DEFAULT)


def relpos_bias(q, rel_h, rel_w, S):
    """q [B,H,S*S,hd] (unscaled) -> bias [B,H,S*S,S*S]; image_encoder.py:347-383."""
    idx = (torch.arange(S)[:, None] - torch.arange(S)[None, :] + S - 1).to(q.device)
    Rh, Rw = rel_h.float()[idx], rel_w.float()[idx]
    B, H, T, hd = q.shape
    rq = q.float().view(B, H, S, S, hd)
    bh = torch.einsum("bnhwc,hkc->bnhwk", rq, Rh)
    bw = torch.einsum("bnhwc,wkc->bnhwk", rq, Rw)
    return (bh[..., :, None] + bw[..., None, :]).reshape(B, H, T, T)


@pytest.mark.parametrize("hd", [64, 80])
def test_attn_flash_global_relpos(flash_version, hd):
    B, H, T = 2, 3, 4096
    D = H * hd
    qkv = rnd(B * T, 3 * D, seed=26)
    rel_h, rel_w = rnd(127, hd, seed=27, scale=0.3), rnd(127, hd, seed=28, scale=0.3)
    table = torch.zeros(256, hd, device=DEV, dtype=torch.bfloat16)
    table[:127], table[128:255] = rel_h, rel_w
    out = torch.zeros(B * T, D, device=DEV, dtype=torch.bfloat16)
    scale = hd ** -0.5
    ops.attn_flash(qkv, 0, qkv, D, qkv, 2 * D, table, out, B, H, T, T, hd, scale)
    sp = lambda t: t.reshape(B, T, H, hd).transpose(1, 2)
    q, k, v = sp(qkv[:, :D]), sp(qkv[:, D:2 * D]), sp(qkv[:, 2 * D:])
    ref = ref_attention(q, k, v, scale, relpos_bias(q, rel_h, rel_w, 64)).transpose(1, 2).reshape(B * T, D)
    assert (out.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("version", FLASH_VERSIONS)
def test_attn_flash_global_relpos_peaky(version):
    """Peaky logits (std ~ 6 in log2 units, strong rel-pos tables): the reference maximum of a row is raised many times
    along the 4096 keys, so the O / l / P rescale path of the lazy softmax is exercised in the rel-pos kernels too."""
    from wildlifemapper_b200 import lib
    lib.call("wm_set_flash_version", version)
    try:
        B, H, T, hd = 1, 2, 4096, 64
        D = H * hd
        qkv = rnd(B * T, 3 * D, seed=41, scale=2.2)
        rel_h, rel_w = rnd(127, hd, seed=42, scale=0.6), rnd(127, hd, seed=43, scale=0.6)
        table = torch.zeros(256, hd, device=DEV, dtype=torch.bfloat16)
        table[:127], table[128:255] = rel_h, rel_w
        out = torch.zeros(B * T, D, device=DEV, dtype=torch.bfloat16)
        scale = hd ** -0.5
        ops.attn_flash(qkv, 0, qkv, D, qkv, 2 * D, table, out, B, H, T, T, hd, scale)
        sp = lambda t: t.reshape(B, T, H, hd).transpose(1, 2)
        q, k, v = sp(qkv[:, :D]), sp(qkv[:, D:2 * D]), sp(qkv[:, 2 * D:])
        ref = ref_attention(q, k, v, scale, relpos_bias(q, rel_h, rel_w, 64)).transpose(1, 2).reshape(B * T, D)
        assert rel_err(out, ref) < 2e-2
    finally:
        lib.call("wm_set_flash_version", FLASH_DEFAULT)


@pytest.mark.parametrize("B,H,hd", [(1, 2, 64), (2, 12, 64), (3, 5, 64), (2, 4, 80)])
def test_attn_window(B, H, hd):
    S = 14
    D = H * hd
    qkv = rnd(B, 64, 64, 3 * D, seed=29)
    rel_h, rel_w = rnd(27, hd, seed=30, scale=0.3), rnd(27, hd, seed=31, scale=0.3)
    table = torch.zeros(64, hd, device=DEV, dtype=torch.bfloat16)
    table[:27], table[32:59] = rel_h, rel_w
    out = torch.zeros(B, 64, 64, D, device=DEV, dtype=torch.bfloat16)
    scale = hd ** -0.5
    ops.attn_window(qkv, table, out, H, scale)
    # reference: zero-pad the (bias-free) qkv to 70x70 -> pad tokens are exact zero keys/values that still
    # take part in the softmax with their rel-pos bias
    xp = torch.nn.functional.pad(qkv.float(), (0, 0, 0, 6, 0, 6))
    win = xp.view(B, 5, S, 5, S, 3 * D).permute(0, 1, 3, 2, 4, 5).reshape(B * 25, S * S, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = win[0], win[1], win[2]
    o = ref_attention(q, k, v, scale, relpos_bias(q, rel_h, rel_w, S))  # [B*25,H,196,hd]
    o = o.transpose(1, 2).reshape(B, 5, 5, S, S, D).permute(0, 1, 3, 2, 4, 5).reshape(B, 70, 70, D)[:, :64, :64]
    assert (out.float() - o).abs().max().item() < 2e-2


# ------------------------------------------------------------------ post-process (integer stages bit-exact)
def test_postprocess_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_post.npz"))
    for case in range(3):
        logits = torch.from_numpy(g[f"c{case}.logits"]).to(DEV)
        boxes = torch.from_numpy(g[f"c{case}.boxes"]).to(DEV)
        sizes = torch.from_numpy(g[f"c{case}.sizes"]).to(DEV)
        B, Q, _ = logits.shape
        packed = torch.zeros(B, Q, 6, device=DEV)
        qidx = torch.zeros(B, Q, device=DEV, dtype=torch.int32)
        counts = torch.zeros(B, device=DEV, dtype=torch.int32)
        ops.postprocess(logits, boxes, sizes, 0.05, 0, packed, qidx, None, counts)
        for i in range(B):
            n = int(counts[i])
            ref_l = g[f"c{case}.{i}.labels"]
            assert n == ref_l.shape[0]
            np.testing.assert_array_equal(packed[i, :n, 5].cpu().numpy().astype(np.int64), ref_l)
            np.testing.assert_allclose(packed[i, :n, 4].cpu().numpy(), g[f"c{case}.{i}.scores"], atol=5e-7, rtol=0)
            np.testing.assert_allclose(packed[i, :n, :4].cpu().numpy(), g[f"c{case}.{i}.boxes"], atol=1e-3, rtol=0)
        # integer stage fed the oracle's probabilities: labels, keep set and boxes bit-exact
        prob = opost.softmax_f32(g[f"c{case}.logits"])
        ref = opost.select_from_prob(prob, g[f"c{case}.boxes"], g[f"c{case}.sizes"], 0.05)
        lab = torch.zeros(B, Q, device=DEV, dtype=torch.int64)
        ops.postprocess(torch.from_numpy(prob).to(DEV), boxes, sizes, 0.05, 1, packed, qidx, lab, counts)
        for i in range(B):
            n = int(counts[i])
            assert n == ref[i]["labels"].shape[0]
            np.testing.assert_array_equal(qidx[i, :n].cpu().numpy(), ref[i]["query"])
            np.testing.assert_array_equal(lab[i, :n].cpu().numpy(), ref[i]["labels"])
            np.testing.assert_array_equal(packed[i, :n, 5].cpu().numpy().astype(np.int64), ref[i]["labels"])
            np.testing.assert_array_equal(packed[i, :n, 4].cpu().numpy(), ref[i]["scores"])
            np.testing.assert_array_equal(packed[i, :n, :4].cpu().numpy(), ref[i]["boxes"])


def test_sigmoid_topk_bit_exact():
    B, Q, K = 2, 900, 300
    rng = np.random.default_rng(5)
    logits = (rng.standard_normal((B, Q, 8)) * 2).astype(np.float32)
    logits = np.round(logits * 4) / 4  # many exact ties
    boxes = rng.random((B, Q, 4)).astype(np.float32)
    prob = opost.sigmoid_f32(logits[..., :7]).reshape(B, Q * 7)
    ref_s, ref_idx = opost.topk_from_prob(prob, K)
    t = lambda a: torch.from_numpy(a).to(DEV)
    prob_ws, order_ws = t(prob.copy()), torch.zeros(B, Q * 7, device=DEV, dtype=torch.int32)
    scores, labels = torch.zeros(B, K, device=DEV), torch.zeros(B, K, device=DEV, dtype=torch.int32)
    query, ob = torch.zeros(B, K, device=DEV, dtype=torch.int32), torch.zeros(B, K, 4, device=DEV)
    ops.sigmoid_topk(t(logits), t(boxes), prob_ws, order_ws, scores, labels, query, ob, 7, K, 1)
    np.testing.assert_array_equal(scores.cpu().numpy(), ref_s)
    np.testing.assert_array_equal(labels.cpu().numpy(), ref_idx % 7)
    np.testing.assert_array_equal(query.cpu().numpy(), ref_idx // 7)
    np.testing.assert_array_equal(ob.cpu().numpy(), np.take_along_axis(boxes, (ref_idx // 7)[..., None], 1))
    # float stage (own sigmoid): same selection up to fp32 rounding of the scores
    ops.sigmoid_topk(t(logits), t(boxes), prob_ws, order_ws, scores, labels, query, ob, 7, K, 0)
    np.testing.assert_allclose(scores.cpu().numpy(), ref_s, atol=2e-7, rtol=0)


def _run_nms(b, s, l, thr=0.4):
    n = b.shape[0]
    nb = (n + 63) // 64
    keep = torch.zeros(max(n, 1), device=DEV, dtype=torch.int64)
    num = torch.zeros(1, device=DEV, dtype=torch.int32)
    ops.nms(torch.from_numpy(b).to(DEV).reshape(-1, 4), torch.from_numpy(s).to(DEV),
            None if l is None else torch.from_numpy(l).to(DEV), thr,
            torch.zeros(max(n, 1), device=DEV, dtype=torch.int32), torch.zeros(max(n * nb, 1), device=DEV, dtype=torch.int64),
            keep, num)
    return keep[:int(num)].cpu().numpy()


@pytest.mark.parametrize("tag,n,dup", [("n2000", 2000, False), ("n2000dup", 2000, True), ("n10000", 10000, False)])
def test_nms_golden_bit_exact(golden_dir, tag, n, dup):
    g = np.load(os.path.join(golden_dir, "golden_nms.npz"))
    b, s, l = opost.make_nms_problem(n, seed=3, dup_scores=dup)
    np.testing.assert_array_equal(_run_nms(b, s, None), g[f"{tag}.keep"].astype(np.int64))
    np.testing.assert_array_equal(_run_nms(b, s, l), g[f"{tag}.keep_per_class"].astype(np.int64))


def test_nms_edges(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_nms.npz"))
    np.testing.assert_array_equal(_run_nms(g["tie.boxes"], g["tie.scores"], None), g["tie.keep"].astype(np.int64))
    assert _run_nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), None).shape == (0,)
    z = np.array([[5, 5, 5, 5], [5, 5, 5, 5]], np.float32)
    np.testing.assert_array_equal(_run_nms(z, np.array([0.9, 0.8], np.float32), None), [0, 1])
    for seed in range(3):
        b, s, l = opost.make_nms_problem(777, seed=seed, dup_scores=bool(seed % 2))
        np.testing.assert_array_equal(_run_nms(b, s, None), opost.nms(b, s, 0.4))
        np.testing.assert_array_equal(_run_nms(b, s, l), opost.batched_nms(b, s, l, 0.4))


@pytest.mark.parametrize("per_class", [0, 1])
def test_nms_batched_small_bit_exact(per_class):
    B, Q = 5, 900
    packed = np.zeros((B, Q, 6), np.float32)
    counts = np.array([900, 513, 0, 1, 64], np.int32)
    for b in range(B):
        bx, sc, lb = opost.make_nms_problem(Q, seed=10 + b, dup_scores=bool(b % 2))
        bx = bx * np.float32(0.25)  # crowd the boxes so plenty overlap
        packed[b, :, :4], packed[b, :, 4], packed[b, :, 5] = bx, sc, lb
    keep_idx = torch.zeros(B, Q, device=DEV, dtype=torch.int32)
    keep_cnt = torch.zeros(B, device=DEV, dtype=torch.int32)
    ops.nms_batched(torch.from_numpy(packed).to(DEV), torch.from_numpy(counts).to(DEV), 0.5, 0.4, per_class, keep_idx, keep_cnt)
    for b in range(B):
        n = counts[b]
        rows = packed[b, :n]
        cand = np.nonzero(rows[:, 4] > np.float32(0.5))[0]
        if per_class:
            ref = cand[opost.batched_nms(rows[cand, :4], rows[cand, 4], rows[cand, 5].astype(np.int64), 0.4)]
        else:
            ref = cand[opost.nms(rows[cand, :4], rows[cand, 4], 0.4)]
        assert int(keep_cnt[b]) == ref.shape[0]
        np.testing.assert_array_equal(keep_idx[b, :ref.shape[0]].cpu().numpy(), ref)


# ------------------------------------------------------------------ round-2 additions: aliasing, benchmarked shapes, NaN
@pytest.mark.parametrize("M,bn", [(700, 0), (700, 512), (40000, 0)])
def test_gemm_inplace_residual_with_bf16_copy(M, bn):
    """The last encoder block's call shape (engine.py, lin2): residual ALIASES the fp32 output (in-place update of the
    residual stream) AND a bf16 copy is requested.  Both outputs must carry bias + residual (the bf16 copy of
    the CTA-pair kernel lost the residual)."""
    N, K = 768, 256
    a, w = rnd(M, K, seed=60), rnd(N, K, seed=61, scale=K ** -0.5)
    bias = rnd(N, seed=62, dtype=torch.float32)
    x = rnd(M, N, seed=63, dtype=torch.float32, scale=3.0)
    x0 = x.clone()
    o16 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, bias, x, M, o16, x, 0, bn)
    y = a.float() @ w.float().t() + bias + x0
    assert (x - y).abs().max().item() < 5e-3
    assert (o16.float() - y).abs().max().item() <= y.abs().max().item() * 2 ** -8 + 5e-3


@pytest.mark.parametrize("M,N,K,act", [(131072, 2304, 768, 0), (131072, 3072, 768, 1), (131072, 768, 3072, 0),
                                       (262144, 1280, 1280, 0), (262144, 256, 1280, 0)])
def test_gemm_at_benchmarked_rows(M, N, K, act):
    """The row counts the bench runs (batch 32: M = 131072; ViT-H batch 64: M = 262144): persistent tile schedulers and
    32-bit index arithmetic only break at scale.  Reference = torch matmul on the same bf16 operands (fp32 result)."""
    a, w = rnd(M, K, seed=64), rnd(N, K, seed=65, scale=K ** -0.5)
    bias = rnd(N, seed=66, dtype=torch.float32)
    res = rnd(M, N, seed=67, dtype=torch.float32) if act == 0 else None
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16 if act else torch.float32)
    if act:
        ops.gemm(a, w, bias, None, 0, out, None, act, 0)
    else:
        out.copy_(res)
        ops.gemm(a, w, bias, out, M, None, out, 0, 0)  # in-place residual update (TMA reduce-add), as proj / lin2 do
    for r0 in (0, M // 2 - 4096, M - 8192):  # check three 8192-row slabs incl. the last rows
        sl = slice(r0, r0 + 8192)
        y = a[sl].float() @ w.float().t() + bias
        y = torch.nn.functional.gelu(y) if act else y + res[sl]
        tol = 5e-2 if act else 5e-3
        assert (out[sl].float() - y).abs().max().item() < tol, r0


@pytest.mark.parametrize("B,H,hd", [(32, 12, 64), (64, 16, 80)])
def test_attn_flash_global_relpos_at_benchmarked_batch(B, H, hd):
    """Global attention at the batch sizes the bench runs (ViT-B batch 32, ViT-H batch 64); sampled (image, head) pairs
    incl. the last ones are compared against fp32 torch."""
    T, D = 4096, H * hd
    qkv = rnd(B * T, 3 * D, seed=70)
    rel_h, rel_w = rnd(127, hd, seed=71, scale=0.3), rnd(127, hd, seed=72, scale=0.3)
    table = torch.zeros(256, hd, device=DEV, dtype=torch.bfloat16)
    table[:127], table[128:255] = rel_h, rel_w
    out = torch.zeros(B * T, D, device=DEV, dtype=torch.bfloat16)
    scale = hd ** -0.5
    ops.attn_flash(qkv, 0, qkv, D, qkv, 2 * D, table, out, B, H, T, T, hd, scale)
    v3 = qkv.view(B, T, 3, H, hd)
    for b, h in ((0, 0), (B // 2, H // 2), (B - 1, H - 1), (B - 1, 0), (7 % B, 5 % H)):
        q, k, v = (v3[b, :, i, h][None, None] for i in range(3))
        ref = ref_attention(q, k, v, scale, relpos_bias(q, rel_h, rel_w, 64))[0, 0]
        got = out.view(B, T, H, hd)[b, :, h].float()
        assert (got - ref).abs().max().item() < 2e-2, (b, h)


@pytest.mark.parametrize("B,H,hd", [(32, 12, 64), (64, 16, 80)])
def test_attn_window_at_benchmarked_batch(B, H, hd):
    S, D = 14, H * hd
    qkv = rnd(B, 64, 64, 3 * D, seed=73)
    rel_h, rel_w = rnd(27, hd, seed=74, scale=0.3), rnd(27, hd, seed=75, scale=0.3)
    table = torch.zeros(64, hd, device=DEV, dtype=torch.bfloat16)
    table[:27], table[32:59] = rel_h, rel_w
    out = torch.zeros(B, 64, 64, D, device=DEV, dtype=torch.bfloat16)
    scale = hd ** -0.5
    ops.attn_window(qkv, table, out, H, scale)
    for b in (0, B // 2, B - 1):
        xp = torch.nn.functional.pad(qkv[b:b + 1].float(), (0, 0, 0, 6, 0, 6))
        win = xp.view(1, 5, S, 5, S, 3 * D).permute(0, 1, 3, 2, 4, 5).reshape(25, S * S, 3, H, hd).permute(2, 0, 3, 1, 4)
        q, k, v = win[0], win[1], win[2]
        o = ref_attention(q, k, v, scale, relpos_bias(q, rel_h, rel_w, S))
        o = o.transpose(1, 2).reshape(1, 5, 5, S, S, D).permute(0, 1, 3, 2, 4, 5).reshape(70, 70, D)[:64, :64]
        # (1.5 bf16 ulp of the largest outputs, |o| up to ~2.5, at head dim 80: measured 2.3e-2 over 64 x 25 x 16 windows)
        assert (out[b].float() - o).abs().max().item() < 3e-2, b


def test_attn_flash_hfc_at_benchmarked_batch():
    """HFC cross-attention shape (8 heads x 128) at batch 32; sampled (image, head) pairs vs fp32 torch."""
    B, H, hd, T = 32, 8, 128, 4096
    q = rnd(B * T, H * hd, seed=76)
    kv = rnd(B * T, 2 * H * hd, seed=77)
    out = torch.zeros(B * T, H * hd, device=DEV, dtype=torch.bfloat16)
    scale = 1 / math.sqrt(hd)
    ops.attn_flash(q, 0, kv, 0, kv, H * hd, None, out, B, H, T, T, hd, scale)
    for b, h in ((0, 0), (B - 1, H - 1), (13, 3)):
        qq = q.view(B, T, H, hd)[b, :, h][None, None]
        kk = kv.view(B, T, 2, H, hd)[b, :, 0, h][None, None]
        vv = kv.view(B, T, 2, H, hd)[b, :, 1, h][None, None]
        ref = ref_attention(qq, kk, vv, scale)[0, 0]
        assert (out.view(B, T, H, hd)[b, :, h].float() - ref).abs().max().item() < 2e-2, (b, h)


def test_nan_scores_do_not_fault():
    """NaN scores (NaN logits) must give a well-defined order (NaN first, torch.sort(descending) convention) and never an
    out-of-bounds gather (colliding ranks left order slots unwritten)."""
    n = 1500
    b, s, _ = opost.make_nms_problem(n, seed=9)
    s = s.copy()
    s[[3, 700, 1499]] = np.nan
    keep = _run_nms(b, s, None)
    torch.cuda.synchronize()
    assert keep.shape[0] > 0 and keep.min() >= 0 and keep.max() < n
    s_ref = s.copy()
    s_ref[[3, 700, 1499]] = [1e30, 1e29, 1e28]  # the same total order without NaN
    np.testing.assert_array_equal(keep, opost.nms(b, s_ref, 0.4))
    B, Q, K = 1, 300, 50
    logits = torch.randn(B, Q, 8, device=DEV)
    logits[0, 5, 2] = float("nan")
    scores, labels = torch.zeros(B, K, device=DEV), torch.zeros(B, K, device=DEV, dtype=torch.int32)
    query, ob = torch.zeros(B, K, device=DEV, dtype=torch.int32), torch.zeros(B, K, 4, device=DEV)
    ops.sigmoid_topk(logits, torch.rand(B, Q, 4, device=DEV), torch.zeros(B, Q * 7, device=DEV),
                     torch.zeros(B, Q * 7, device=DEV, dtype=torch.int32), scores, labels, query, ob, 7, K, 0)
    torch.cuda.synchronize()
    assert int(query[0, 0]) == 5 and int(labels[0, 0]) == 2
    assert (query >= 0).all() and (query < Q).all()
