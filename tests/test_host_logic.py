"""CPU tests of the host-side logic: drop-in module surface / state_dict contract, weight-preparation algebra
(bias folding for pad keys, low-pass operator), CPU tensors rejected loudly, tile sharding + gloo gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wildlifemapper_b200"))

import segment_anything as sa  # noqa: E402
from segment_anything.network import MedSAM  # noqa: E402
from segment_anything.utils.misc import NestedTensor, nested_tensor_from_tensor_list  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle.weights import make_state_dict, state_dict_spec  # noqa: E402
from wildlifemapper_b200.dist import shard_bounds  # noqa: E402
from wildlifemapper_b200.engine import EncoderEngine, lowpass_operator_tables  # noqa: E402


@pytest.mark.parametrize("mt", ["vit_b", "vit_l", "vit_h"])
def test_dropin_state_dict_contract(mt):
    sam, crit, post = sa.sam_model_registry[mt]()
    m = MedSAM(sam.image_encoder, sam.mask_decoder, sam.prompt_encoder)
    spec = state_dict_spec(mt)
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec.keys())  # silent strict=False renames are the failure mode
    assert all(tuple(sd[k].shape) == spec[k] for k in sd)
    m.load_state_dict(make_state_dict(mt) if mt == "vit_b" else sd, strict=True)
    # freezing policy of MedSAM (network.py:19-34)
    names = {n for n, p in m.named_parameters() if p.requires_grad}
    assert any("hfc_attn" in n for n in names) and any("patch_embed" in n for n in names)
    assert not any(".blocks." in n for n in names)
    assert "bbox" in post and hasattr(sam.image_encoder, "img_size")


def test_import_surface():
    from segment_anything import SamPredictor, build_sam, build_sam_vit_b, build_sam_vit_h, build_sam_vit_l  # noqa: F401
    from segment_anything.modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, Sam, TwoWayTransformer  # noqa: F401
    from segment_anything.utils import misc
    for n in ("custom_collate", "MetricLogger", "SmoothedValue", "reduce_dict", "all_gather", "NestedTensor"):
        assert hasattr(misc, n)
    from segment_anything.utils.augmentation_yolo import random_perspective  # noqa: F401
    args = type("A", (), {"num_queries": 900, "device": "cpu"})()
    sam, _, _ = sa.sam_model_registry["vit_b"](args=args)
    assert sam.mask_decoder.mask_tokens.weight.shape == (900, 256)


def test_cpu_tensors_are_rejected_not_emulated():
    sam, _, _ = sa.sam_model_registry["vit_b"]()
    m = MedSAM(sam.image_encoder, sam.mask_decoder, sam.prompt_encoder).eval()
    nt = nested_tensor_from_tensor_list([torch.zeros(3, 700, 900)])
    assert tuple(nt.tensors.shape) == (1, 3, 1024, 1024) and not nt.mask[0, :700, :900].any()
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(NestedTensor(nt.tensors, None), None)


def test_lowpass_tables_match_oracle_operator():
    Lr, Li = lowpass_operator_tables()
    Or, Oi = om.lowpass_operator()
    assert (Lr - Or).abs().max() < 1e-12 and (Li - Oi).abs().max() < 1e-12


def test_bias_folding_is_exact_algebra():
    """Dropping the k/v biases and folding W_proj b_v into the proj bias (engine.prepare) leaves the window block
    unchanged, INCLUDING the pad keys that take part in the softmax."""
    sd = make_state_dict("vit_t", seed=3)
    pre = "image_encoder.blocks.0."
    D = 128
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 64, 64, D, generator=g)
    ref = om.block(x, sd, pre, heads=2, window=14)
    sd2 = dict(sd)
    qb = sd[pre + "attn.qkv.bias"].clone()
    bv = qb[2 * D:].clone()
    qb[D:] = 0
    sd2[pre + "attn.qkv.bias"] = qb
    sd2[pre + "attn.proj.bias"] = sd[pre + "attn.proj.bias"] + sd[pre + "attn.proj.weight"] @ bv
    # folded parameters + zero-padding of the qkv tensor itself (what the TMA out-of-bounds fill produces)
    xn = F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-6)
    qkv = F.linear(xn, sd2[pre + "attn.qkv.weight"], qb)
    qkvp = F.pad(qkv, (0, 0, 0, 6, 0, 6)).view(1, 5, 14, 5, 14, 3, 2, 64).permute(5, 0, 1, 3, 6, 2, 4, 7).reshape(3, 25, 2, 196, 64)
    q, k, v = qkvp[0], qkvp[1], qkvp[2]
    idx = torch.arange(14)[:, None] - torch.arange(14)[None, :] + 13
    Rh, Rw = sd[pre + "attn.rel_pos_h"][idx], sd[pre + "attn.rel_pos_w"][idx]
    rq = q.reshape(25, 2, 14, 14, 64)
    bias = (torch.einsum("bnhwc,hkc->bnhwk", rq, Rh)[..., :, None] + torch.einsum("bnhwc,wkc->bnhwk", rq, Rw)[..., None, :])
    s = (q * 64 ** -0.5) @ k.transpose(-1, -2) + bias.reshape(25, 2, 196, 196)
    o = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(1, 5, 5, 14, 14, D).permute(0, 1, 3, 2, 4, 5).reshape(1, 70, 70, D)[:, :64, :64]
    y = x + F.linear(o, sd2[pre + "attn.proj.weight"], sd2[pre + "attn.proj.bias"])
    xn2 = F.layer_norm(y, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-6)
    y = y + F.linear(F.gelu(F.linear(xn2, sd[pre + "mlp.lin1.weight"], sd[pre + "mlp.lin1.bias"])),
                     sd[pre + "mlp.lin2.weight"], sd[pre + "mlp.lin2.bias"])
    assert (y - ref).abs().max() < 2e-5


def test_engine_prepare_layouts_on_cpu():
    """Weight preparation is pure layout work and runs on any device."""
    sd = {k[len("image_encoder."):]: v for k, v in make_state_dict("vit_t", seed=0).items() if k.startswith("image_encoder.")}
    eng = EncoderEngine(128, 2, 2, (1,), device="cpu")
    eng.prepare(sd)
    w = eng.w
    assert w["b0.rel"].shape == (64, 64) and w["b1.rel"].shape == (256, 64)
    assert torch.equal(w["b0.rel"][:27].float(), sd["blocks.0.attn.rel_pos_h"].to(torch.bfloat16).float())
    assert torch.equal(w["b1.rel"][128:255].float(), sd["blocks.1.attn.rel_pos_w"].to(torch.bfloat16).float())
    assert w["b0.qkv_b"][128:].abs().max() == 0 and w["lp1"].shape == (2048, 1024) and w["lp2"].shape == (1024, 2048)
    # conv3x3 weight layout: k = (dy*3+dx)*C + c
    wt = sd["neck.2.weight"]
    assert torch.equal(w["neck2_w"][5, (1 * 3 + 2) * 256 + 7].float(), wt[5, 7, 1, 2].to(torch.bfloat16).float())
    # ViT-H (head dim 80) is supported: its rel-pos tables are [*, 80]; any other head dim is a loud error
    sd80 = {k[len("image_encoder."):]: v for k, v in make_state_dict("vit_t80", seed=0).items() if k.startswith("image_encoder.")}
    eng80 = EncoderEngine(640, 2, 8, (1,), device="cpu")
    eng80.prepare(sd80)
    assert eng80.w["b0.rel"].shape == (64, 80) and eng80.w["b1.rel"].shape == (256, 80)
    with pytest.raises(NotImplementedError):
        EncoderEngine(768, 12, 8, (2,), device="cpu")  # head dim 96


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 32, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wildlifemapper_b200.dist import gather_detections, shard_bounds
    lo, hi = shard_bounds(6, rank, world)
    allp = torch.arange(6 * 4 * 6, dtype=torch.float32).view(6, 4, 6)
    allc = torch.tensor([1, 0, 4, 2, 3, 1], dtype=torch.int32)
    p, c, k, kc = gather_detections(allp[lo:hi].clone(), allc[lo:hi].clone(), None, None)
    # packed variant: the four detection tensors of a rank in one flat buffer, ONE all-gather, split on arrival
    from wildlifemapper_b200.dist import gather_buffer
    from wildlifemapper_b200.postprocess import DetectionBuffer
    B, Q = 3, 4
    allk = (torch.arange(6 * Q, dtype=torch.int32) * 7 % 5).view(6, Q)
    allkc = torch.tensor([0, 4, 1, 2, 2, 3], dtype=torch.int32)
    buf = DetectionBuffer(B, Q, "cpu")
    buf.packed.copy_(allp[lo:hi]); buf.counts.copy_(allc[lo:hi]); buf.keep_idx.copy_(allk[lo:hi]); buf.keep_cnt.copy_(allkc[lo:hi])
    gp, gc, gk, gkc = DetectionBuffer.split(gather_buffer(buf), world, B, Q)
    ok_buf = torch.equal(gp, allp) and torch.equal(gc, allc) and torch.equal(gk, allk) and torch.equal(gkc, allkc)
    q.put((rank, torch.equal(p, allp), torch.equal(c, allc), k is None and ok_buf))
    dist.destroy_process_group()


def test_gather_detections_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res)


def test_plan_tiles_covers_image_with_overlap():
    """Host logic of the survey front-end (wildlifemapper_b200/survey.py): every pixel is covered, neighbours share at
    least `overlap` pixels, no tile hangs over the border unless the image is smaller than a tile."""
    from wildlifemapper_b200.survey import plan_tiles
    for H, W, tile, ov in [(3648, 5472, 1024, 128), (1024, 1024, 1024, 128), (700, 2000, 1024, 0), (1025, 1024, 1024, 512),
                           (4000, 3000, 768, 64)]:
        org = plan_tiles(H, W, tile, ov)
        cover = np.zeros((H, W), bool)
        for y, x in org:
            assert y >= 0 and x >= 0 and (y + tile <= H or H <= tile) and (x + tile <= W or W <= tile)
            cover[y:y + tile, x:x + tile] = True
        assert cover.all() and len(set(org)) == len(org)
        ys, xs = sorted({y for y, _ in org}), sorted({x for _, x in org})
        assert all(b - a <= tile - ov for a, b in zip(ys, ys[1:])) and all(b - a <= tile - ov for a, b in zip(xs, xs[1:]))
    assert plan_tiles(3648, 5472)[-1] == (3648 - 1024, 5472 - 1024) and len(plan_tiles(3648, 5472)) == 24
    with pytest.raises(ValueError):
        plan_tiles(0, 10)
    with pytest.raises(ValueError):
        plan_tiles(10, 10, 1024, 1024)


def test_loader_transforms_match_reference_goldens(golden_dir):
    """The drop-in ``segment_anything.utils.augmentation`` (what dataloader_coco.py:275-292 composes) against goldens minted by
    running the reference's own transform classes: image tensors sha256-exact, targets exact."""
    import hashlib
    import random
    import sys as _sys
    from PIL import Image
    _sys.path.insert(0, os.path.join(ROOT, "wildlifemapper_b200"))
    import segment_anything.utils.augmentation as T
    from oracle.frontend import frontend_image
    g = np.load(os.path.join(golden_dir, "golden_frontend.npz"))
    for tag, hw in (("pipe_sq", (1024, 1024)), ("pipe_wide", (600, 900))):
        img = frontend_image(tag, hw)
        gen = torch.Generator().manual_seed(11)
        xy = torch.rand(9, 2, generator=gen) * torch.tensor([hw[1] - 80.0, hw[0] - 80.0])
        wh = torch.rand(9, 2, generator=gen) * 60 + 8
        tgt = {"boxes": torch.cat([xy, xy + wh], 1), "area": wh[:, 0] * wh[:, 1], "center": xy + wh / 2,
               "labels": torch.arange(9) % 6 + 1, "orig_size": torch.as_tensor([hw[0], hw[1]]), "size": torch.as_tensor([hw[0], hw[1]])}
        norm = [T.RandomResize([768], max_size=768), T.ToTensor(), T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])]
        for mode, tf in (("val", norm), ("train", norm + [T.FlipLR(fliplr=1.0)])):
            random.seed(0)
            im, tg = T.Compose(tf)(Image.fromarray(img), {k: v.clone() for k, v in tgt.items()})
            a = im.numpy()
            assert tuple(a.shape) == tuple(g[f"{tag}.{mode}.shape"])
            assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == str(g[f"{tag}.{mode}.sha256"])
            for k in ("boxes", "area", "center", "size"):
                np.testing.assert_array_equal(tg[k].numpy(), g[f"{tag}.{mode}.{k}"], err_msg=f"{tag}.{mode}.{k}")
    with pytest.raises(NotImplementedError):
        T.RandomSizeCrop(384, 600)
