"""Pin the CPU oracle (oracle/model.py) against golden vectors minted from the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import model as om
from oracle.weights import make_state_dict, make_tiles, state_dict_spec

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import sample_positions  # noqa: E402


def _check_taps(g, taps, names, tol):
    for name, t in names.items():
        f = taps[t].detach().float().contiguous().view(-1)
        pos = sample_positions(f.numel(), name)
        got = f[torch.from_numpy(pos)].numpy()
        ref = g[f"{name}.samples"]
        assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max()), name
        stats = g[f"{name}.stats"]
        assert abs(f.abs().mean().item() - stats[1]) <= tol * max(1.0, stats[1]), name


@pytest.mark.parametrize("tag,batch,nq", [("vit_t", 2, 51), ("vit_t_q900", 1, 900)])
def test_oracle_matches_reference_tiny(golden_dir, tag, batch, nq):
    g = np.load(os.path.join(golden_dir, f"golden_model_{tag}.npz"))
    sd = make_state_dict("vit_t", seed=0, num_queries=nq)
    taps = {}
    out = om.forward(sd, "vit_t", make_tiles(batch, seed=2), taps)
    # fp32 tolerance: the oracle restates the same arithmetic with the same ATen kernels
    assert np.abs(out["pred_logits"].numpy() - g["pred_logits"]).max() < 2e-4
    assert np.abs(out["pred_boxes"].numpy() - g["pred_boxes"]).max() < 2e-5
    names = {"x_hfc": "x_hfc", "hfc_tok": "hfc_tok", "block0": "block0", "block1": "block1",
             "features": "features", "hs": "hs"}
    _check_taps(g, taps, names, 2e-4)
    # hfc_attn_out is the module output before the residual add
    f = (taps["after_hfc"] - taps["patch_pos"]).contiguous().view(-1)
    pos = sample_positions(f.numel(), "hfc_attn_out")
    assert np.abs(f[torch.from_numpy(pos)].numpy() - g["hfc_attn_out.samples"]).max() < 2e-4


def test_oracle_matches_reference_vit_b_heads(golden_dir):
    """Full ViT-B forward takes ~10 s of CPU; checks logits/boxes + encoder feature samples."""
    path = os.path.join(golden_dir, "golden_model_vit_b.npz")
    g = np.load(path)
    sd = make_state_dict("vit_b", seed=0)
    taps = {}
    out = om.forward(sd, "vit_b", make_tiles(1, seed=2), taps)
    assert np.abs(out["pred_logits"].numpy() - g["pred_logits"]).max() < 5e-4
    assert np.abs(out["pred_boxes"].numpy() - g["pred_boxes"]).max() < 5e-5
    _check_taps(g, taps, {"features": "features", "block11": "block11", "block2": "block2"}, 5e-4)


def test_state_dict_contract():
    """Key set, order and count of the reference state_dict (295 tensors for ViT-B)."""
    spec = state_dict_spec("vit_b")
    assert len(spec) == 295
    assert spec["image_encoder.blocks.2.attn.rel_pos_h"] == (127, 64)
    assert spec["image_encoder.blocks.0.attn.rel_pos_h"] == (27, 64)
    assert state_dict_spec("vit_h")["image_encoder.blocks.7.attn.rel_pos_w"] == (127, 80)


def test_lowpass_operator_equals_fft():
    """App. A.1: x_hfc = |g - Re(L g L^T)| with L = Lr + i Li."""
    torch.manual_seed(0)
    img = torch.randn(1, 3, 1024, 1024)
    ref = om.hfc_highpass(img)
    Lr, Li = om.lowpass_operator()
    g = (0.2989 * img[:, 0] + 0.587 * img[:, 1] + 0.114 * img[:, 2])[0].double()
    low = Lr @ g @ Lr.t() - Li @ g @ Li.t()
    assert (ref[0, 0].double() - (g - low).abs()).abs().max() < 1e-5
