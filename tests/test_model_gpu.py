"""End-to-end parity on the B200 through the drop-in ``segment_anything`` modules (which call the C ABI):
the CUDA path vs (a) golden outputs minted from the unmodified reference and (b) the fp32 CPU oracle run on the
same seeded weights and tiles.  Tolerances are the ones BASELINE.json / SURVEY.md section 8d state for bf16:
logits max-abs <= 2e-2, boxes mean-L1 <= 1e-3 and max <= 5e-3 (normalised), encoder features reported against
the 5e-2 / 6e-3 (max / mean) bf16 proxy."""
import os
import sys
from functools import partial

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wildlifemapper_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import segment_anything as sa  # noqa: E402  (the drop-in)
from segment_anything.modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, TwoWayTransformer  # noqa: E402
from segment_anything.network import MedSAM  # noqa: E402
from segment_anything.utils.misc import NestedTensor  # noqa: E402
from make_golden import sample_positions  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle.weights import MODEL_CONFIGS, make_state_dict, make_tiles  # noqa: E402

DEV = "cuda"


def build(model_type: str, num_queries: int = 51) -> MedSAM:
    D, depth, heads, glob = MODEL_CONFIGS[model_type]
    enc = ImageEncoderViT(depth=depth, embed_dim=D, img_size=1024, mlp_ratio=4,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=heads, patch_size=16,
                          qkv_bias=True, use_rel_pos=True, global_attn_indexes=list(glob), window_size=14, out_chans=256)
    pe = PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16)
    dec = MaskDecoder(num_multimask_outputs=num_queries - 1,
                      transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                      transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256)
    m = MedSAM(image_encoder=enc, mask_decoder=dec, prompt_encoder=pe).eval()
    m.load_state_dict(make_state_dict(model_type, seed=0, num_queries=num_queries), strict=True)
    return m.to(DEV)


def check_outputs(out, ref_logits, ref_boxes, tag):
    dl = np.abs(out["pred_logits"].cpu().numpy() - ref_logits)
    db = np.abs(out["pred_boxes"].cpu().numpy() - ref_boxes)
    print(f"[{tag}] logits max-abs {dl.max():.3e}  boxes mean-L1 {db.mean():.3e} max {db.max():.3e}")
    assert out["pred_logits"].dtype == torch.float32 and out["pred_boxes"].dtype == torch.float32
    assert dl.max() <= 2e-2
    assert db.mean() <= 1e-3 and db.max() <= 5e-3


@pytest.mark.parametrize("tag,batch,nq", [("vit_t", 2, 51), ("vit_t_q900", 1, 900)])
def test_tiny_model_vs_reference_golden(golden_dir, tag, batch, nq):
    g = np.load(os.path.join(golden_dir, f"golden_model_{tag}.npz"))
    model = build("vit_t", nq)
    tiles = make_tiles(batch, seed=2).to(DEV)
    with torch.no_grad():
        out = model(NestedTensor(tiles, None), np.array([[0, 0, 1024, 1024]] * batch))
    check_outputs(out, g["pred_logits"], g["pred_boxes"], tag)


def test_head_dim_80_model_vs_oracle():
    """ViT-H's head dim (80) on a 2-block encoder (one windowed, one global block): the sm_100a path against the fp32
    CPU oracle on the same seeded weights and tiles (the oracle itself is pinned to the reference for vit_t / vit_b)."""
    model = build("vit_t80", 51)
    tiles = make_tiles(2, seed=2)
    with torch.no_grad():
        out = model(NestedTensor(tiles.to(DEV), None), None)
    ref = om.forward(make_state_dict("vit_t80", seed=0), "vit_t80", tiles)
    check_outputs(out, ref["pred_logits"].numpy(), ref["pred_boxes"].numpy(), "vit_t80")


def test_vit_b_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_model_vit_b.npz"))
    sam, _crit, post = sa.sam_model_registry["vit_b"]()
    model = MedSAM(sam.image_encoder, sam.mask_decoder, sam.prompt_encoder).eval()
    model.load_state_dict(make_state_dict("vit_b", seed=0), strict=True)
    model = model.to(DEV)
    tiles = make_tiles(1, seed=2).to(DEV)
    with torch.no_grad():
        mask = model.fft(NestedTensor(tiles, None))
        feats = model.image_encoder(tiles, mask)
        out = model(NestedTensor(tiles, None), None)
    check_outputs(out, g["pred_logits"], g["pred_boxes"], "vit_b")
    # encoder features [B,256,64,64] at the golden sample positions
    f = feats.float().contiguous().view(-1)
    got = f[torch.from_numpy(sample_positions(f.numel(), "features")).to(DEV)].cpu().numpy()
    d = np.abs(got - g["features.samples"])
    print(f"[vit_b] features max-abs {d.max():.3e} mean-abs {d.mean():.3e} (|ref| max {np.abs(g['features.samples']).max():.2f})")
    assert d.max() <= 5e-2 and d.mean() <= 6e-3  # the bf16 proxy of SURVEY.md section 0.7 (measured 1.9e-2 / 4.7e-3)
    # high-pass image vs the reference's FFT path
    h = mask.float().contiguous().view(-1)
    got = h[torch.from_numpy(sample_positions(h.numel(), "x_hfc")).to(DEV)].cpu().numpy()
    assert np.abs(got - g["x_hfc.samples"]).max() <= 2e-4  # split low-pass products (single bf16 operands: 6e-3)
    # drop-in PostProcess on the CUDA outputs vs the reference PostProcess on the reference outputs
    res = post["bbox"](out, torch.tensor([[1024, 1024]], device=DEV))
    assert res[0]["labels"].dtype == torch.int64
    n_ref = g["pp0.labels"].shape[0]
    assert abs(res[0]["labels"].shape[0] - n_ref) <= 2  # scores near the 0.05 threshold may flip in bf16
    if res[0]["labels"].shape[0] == n_ref:
        assert (res[0]["labels"].cpu().numpy() == g["pp0.labels"]).mean() >= 0.9
        assert np.abs(res[0]["boxes"].cpu().numpy() - g["pp0.boxes"]).max() <= 5e-3 * 1024


def test_highpass_on_smooth_imagery():
    """x_hfc = |gray - lowpass| is a small difference of O(1) numbers on smooth (natural-image-like) tiles; with
    single bf16 operands in the DFT-operator GEMMs the error (1.4e-3 mean) is 5 % of the signal on a 1/f^1.5 field.  The
    split (hi + lo) products must reproduce the reference's fp32 FFT path (network.py:43-55) to 1e-3 of the signal."""
    gen = torch.Generator().manual_seed(5)
    fy, fx = torch.fft.fftfreq(1024)[:, None], torch.fft.fftfreq(1024)[None, :]
    f = torch.sqrt(fy ** 2 + fx ** 2)
    f[0, 0] = 1.0
    tiles = []
    for alpha in (1.0, 1.5):
        spec = torch.complex(torch.randn(1024, 1024, generator=gen), torch.randn(1024, 1024, generator=gen)) / f ** alpha
        field = torch.fft.ifft2(spec).real
        field = (field - field.mean()) / field.std()
        tiles.append(torch.stack([field + 0.3, 0.8 * field - 0.1, 1.1 * field + 0.5]))
    tiles = torch.stack(tiles).float().contiguous()
    ref = om.hfc_highpass(tiles)
    model = build("vit_t", 51)
    eng = model.image_encoder.engine()
    assert eng.hfc_precision == "split"
    with torch.no_grad():
        _, a_hfc, hfc_img = eng.highpass(tiles.to(DEV), want_image=True)
    torch.cuda.synchronize()
    for i, alpha in enumerate((1.0, 1.5)):
        d = (hfc_img[i].cpu() - ref[i]).abs()
        rel = d.mean().item() / ref[i].mean().item()
        print(f"[highpass 1/f^{alpha}] mean |x_hfc| {ref[i].mean():.4f} mean err {d.mean():.2e} max err {d.max():.2e} rel {rel:.2e}")
        assert rel <= 1e-3 and d.max().item() <= 2e-4
    # the im2col rows of hfc_embed are the bf16 rounding of the same image
    rows = torch.nn.functional.unfold(ref, 16, stride=16).transpose(1, 2).reshape(-1, 256)
    assert (a_hfc.float().cpu() - rows).abs().max().item() <= rows.abs().max().item() * 2 ** -8


def test_highpass_batch_independence_and_zero_image():
    """Size-independent properties of E0: a tile's high-pass image does not depend on its batch neighbours (bitwise), a constant
    tile has no high-frequency content beyond rounding, and both precision modes agree to the single-operand error."""
    model = build("vit_t", 51)
    eng = model.image_encoder.engine()
    tiles = make_tiles(3, seed=9).to(DEV)
    with torch.no_grad():
        _, rows3, img3 = eng.highpass(tiles, want_image=True)
        img3, rows3 = img3.clone(), rows3.clone()
        _, rows1, img1 = eng.highpass(tiles[1:2].contiguous(), want_image=True)
        assert torch.equal(img1[0], img3[1]) and torch.equal(rows1, rows3[4096:8192])
        const = torch.full((1, 3, 1024, 1024), 0.75, device=DEV)
        _, _, imgc = eng.highpass(const, want_image=True)
        assert imgc.abs().max().item() <= 2e-5  # the reference's fp32 FFT leaves ~1e-7 here
    torch.cuda.synchronize()


def test_tiny_model_stage_parity_vs_oracle():
    """Per-stage comparison against the CPU oracle on the same inputs (reports where error enters)."""
    model = build("vit_t", 51)
    sd = make_state_dict("vit_t", seed=0)
    tiles = make_tiles(2, seed=2)
    otaps = {}
    oout = om.forward(sd, "vit_t", tiles, otaps)
    eng = model.image_encoder.engine()
    t = tiles.to(DEV)
    taps = {}
    with torch.no_grad():
        a_patch, a_hfc, hfc_img = eng.highpass(t, want_image=True)
        feat, featb = eng.encode(a_patch, a_hfc, 2, taps)
        out = model(NestedTensor(t, None), None)
    torch.cuda.synchronize()

    def cmp(name, got, ref, tol_max):
        d = (got.float().cpu().reshape(ref.shape) - ref).abs()
        print(f"[stage {name}] max-abs {d.max():.3e} mean-abs {d.mean():.3e} |ref|max {ref.abs().max():.2f}")
        assert d.max().item() <= tol_max, name

    cmp("x_hfc", hfc_img, otaps["x_hfc"], 2e-4)
    cmp("after_hfc", taps["after_hfc"], otaps["after_hfc"], 6e-2)  # (measured 3.1e-2 ... 3.9e-2 on |x| <= 10)
    cmp("block0", taps["block0"], otaps["block0"], 6e-2)
    cmp("block1", taps["block1"], otaps["block1"], 6e-2)
    cmp("features", feat.view(2, 4096, 256).transpose(1, 2), otaps["features"].flatten(2), 6e-2)
    cmp("logits", out["pred_logits"], oout["pred_logits"], 2e-2)
    cmp("boxes", out["pred_boxes"], oout["pred_boxes"], 5e-3)


def test_loader_like_tiles_and_determinism():
    """768x768 content zero-padded into the tile (reference loader, utils/misc.py:50-64) + bitwise repeatability."""
    model = build("vit_t", 51)
    sd = make_state_dict("vit_t", seed=0)
    tiles = make_tiles(1, seed=5, loader_like=True)
    ref = om.forward(sd, "vit_t", tiles)
    with torch.no_grad():
        o1 = model(NestedTensor(tiles.to(DEV), None), None)
        l1, b1 = o1["pred_logits"].clone(), o1["pred_boxes"].clone()
        o2 = model(NestedTensor(tiles.to(DEV), None), None)
    check_outputs({"pred_logits": l1, "pred_boxes": b1}, ref["pred_logits"].numpy(), ref["pred_boxes"].numpy(), "loader")
    assert torch.equal(l1, o2["pred_logits"]) and torch.equal(b1, o2["pred_boxes"])


def test_errors_are_loud():
    model = build("vit_t", 51)
    tiles = make_tiles(1, seed=2)
    with pytest.raises(RuntimeError):  # CPU tensors: no fallback
        with torch.no_grad():
            model(NestedTensor(tiles, None), None)
    t = tiles.to(DEV).requires_grad_(True)
    with pytest.raises(RuntimeError):  # training path not implemented
        model(NestedTensor(t, None), None)


def test_cuda_graph_replay_matches_eager():
    """The captured step (wildlifemapper_b200/graph.py) replays exactly the eager kernels: bit-identical detections,
    also for a second batch fed through the static input."""
    from wildlifemapper_b200 import postprocess as pp
    from wildlifemapper_b200.graph import GraphedDetector
    model = build("vit_t", 51)
    sizes = torch.tensor([[1024, 1024]] * 2, device=DEV)
    det = GraphedDetector(model, batch=2, conf_thr=0.05, nms_score_thr=0.05, iou_thr=0.4)
    for seed in (2, 5):
        tiles = make_tiles(2, seed=seed).to(DEV)
        with torch.no_grad():
            out = model(NestedTensor(tiles, None), None)
            packed, _l, _q, counts = pp.postprocess_packed(out["pred_logits"], out["pred_boxes"], sizes, 0.05)
            keep_idx, keep_cnt = pp.nms_packed(packed, counts, score_thr=0.05, iou_threshold=0.4)
        g_packed, g_counts, g_keep_idx, g_keep_cnt = det(tiles)
        torch.cuda.synchronize()
        assert torch.equal(g_counts, counts) and torch.equal(g_keep_cnt, keep_cnt)
        for b in range(2):
            n, k = int(counts[b]), int(keep_cnt[b])
            assert torch.equal(g_packed[b, :n], packed[b, :n])
            assert torch.equal(g_keep_idx[b, :k], keep_idx[b, :k])


# ------------------------------------------------------------------ round 2: call orders the reference allows
def _decode(model, emb):
    return model.mask_decoder(image_embeddings=emb, image_pe=model.prompt_encoder.get_dense_pe(),
                              sparse_prompt_embeddings=None, dense_prompt_embeddings=None, multimask_output=False,
                              hfc_embed=None)


def test_encode_a_encode_b_decode_a():
    """The encoder output carries a token-major twin that lives in the engine workspace; a later encode overwrites that
    workspace.  Decoding the FIRST embedding afterwards must still decode the first image (it decoded B)."""
    model = build("vit_t", 51)
    A, Bt = make_tiles(1, seed=2).to(DEV), make_tiles(1, seed=7).to(DEV)
    with torch.no_grad():
        ref = model(NestedTensor(A, None), None)
        ref_l, ref_b = ref["pred_logits"].clone(), ref["pred_boxes"].clone()
        emb_a = model.image_encoder(A, model.fft(A))
        emb_b = model.image_encoder(Bt, model.fft(Bt))
        out_a = _decode(model, emb_a)          # twin of emb_a is stale -> transpose path
        la, ba = out_a["pred_logits"].clone(), out_a["pred_boxes"].clone()
        out_b = _decode(model, emb_b)          # twin of emb_b is current
        ref_bt = model(NestedTensor(Bt, None), None)
    assert torch.equal(la, ref_l) and torch.equal(ba, ref_b)
    assert torch.equal(out_b["pred_logits"], ref_bt["pred_logits"]) and torch.equal(out_b["pred_boxes"], ref_bt["pred_boxes"])
    assert not torch.equal(la, out_b["pred_logits"])


def test_inplace_edit_of_the_embedding_is_seen():
    model = build("vit_t", 51)
    A = make_tiles(1, seed=2).to(DEV)
    with torch.no_grad():
        emb = model.image_encoder(A, model.fft(A))
        base = _decode(model, emb)["pred_logits"].clone()
        emb.mul_(0.5)                                      # bumps emb._version: the workspace twin no longer describes it
        got = _decode(model, emb)["pred_logits"].clone()
        want = _decode(model, emb.clone())["pred_logits"]  # a clone has no twin: plain transpose path
    assert torch.equal(got, want)
    assert not torch.equal(got, base)


def test_fft_a_fft_b_encode_a():
    """MedSAM.fft leaves im2col rows for the encoder in the workspace; fft(A); fft(B); image_encoder(A, mask_A) must
    not pick up B's rows, and an in-place edit of the mask must be honoured."""
    model = build("vit_t", 51)
    A, Bt = make_tiles(1, seed=2).to(DEV), make_tiles(1, seed=7).to(DEV)
    with torch.no_grad():
        want = model.image_encoder(A, model.fft(A)).clone()
        mask_a = model.fft(A)
        mask_b = model.fft(Bt)
        got = model.image_encoder(A, mask_a).clone()
        want_b = model.image_encoder(Bt, mask_b).clone()
        mask_a2 = model.fft(A)
        mask_a2.zero_()
        got_zero = model.image_encoder(A, mask_a2).clone()
        want_zero = model.image_encoder(A, torch.zeros_like(mask_a2)).clone()
    assert torch.equal(got, want)
    assert not torch.equal(got, want_b)
    assert torch.equal(got_zero, want_zero) and not torch.equal(got_zero, want)


# ------------------------------------------------------------------ round 2: parity at the benchmarked shapes
def test_vit_b_batch32_matches_single_tile_and_oracle():
    """The bench runs ViT-B at batch 32 (M = 131072 token rows: CTA-pair GEMMs, 6144-CTA attention grids).  Tile i of
    the batch must equal the same tile run alone, and sampled tiles must meet the oracle gates."""
    model = build("vit_b", 51)
    tiles = make_tiles(32, seed=2)
    with torch.no_grad():
        out = model(NestedTensor(tiles.to(DEV), None), None)
        L, Bx = out["pred_logits"].clone(), out["pred_boxes"].clone()
        for i in (0, 13, 31):
            o1 = model(NestedTensor(tiles[i:i + 1].to(DEV), None), None)
            dl = (o1["pred_logits"][0] - L[i]).abs().max().item()
            db = (o1["pred_boxes"][0] - Bx[i]).abs().max().item()
            print(f"[vit_b b32 tile {i} vs alone] logits {dl:.2e} boxes {db:.2e}")
            # different GEMM tilings (CTA-pair vs single-CTA kernel) accumulate in a different order: not bitwise
            assert dl <= 4e-3 and db <= 1e-3, (i, dl, db)
    sd = make_state_dict("vit_b", seed=0)
    for i in (5, 31):
        ref = om.forward(sd, "vit_b", tiles[i:i + 1])
        check_outputs({"pred_logits": L[i:i + 1], "pred_boxes": Bx[i:i + 1]}, ref["pred_logits"].numpy(),
                      ref["pred_boxes"].numpy(), f"vit_b b32 tile {i}")


@pytest.mark.parametrize("model_type", ["vit_l", "vit_h"])
def test_full_size_large_models_vs_oracle(model_type):
    """One full-size ViT-L / ViT-H tile (24 / 32 blocks; head dim 64 / 80) against the fp32 CPU oracle."""
    model = build(model_type, 51)
    tiles = make_tiles(2, seed=2)
    with torch.no_grad():
        out = model(NestedTensor(tiles.to(DEV), None), None)
    ref = om.forward(make_state_dict(model_type, seed=0), model_type, tiles[1:2])
    check_outputs({"pred_logits": out["pred_logits"][1:2], "pred_boxes": out["pred_boxes"][1:2]},
                  ref["pred_logits"].numpy(), ref["pred_boxes"].numpy(), model_type)
