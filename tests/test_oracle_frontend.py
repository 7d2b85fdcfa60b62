"""Pin oracle/frontend.py against the golden minted from the reference's own transform classes and
``nested_tensor_from_tensor_list`` (tests/golden/make_golden.py frontend).  CPU only."""
import hashlib
import os
import zlib

import numpy as np
import pytest

from oracle import frontend as of
from oracle import post as op


def _sample_positions(numel, key):  # as tests/golden/make_golden.py
    return np.random.default_rng(zlib.crc32(key.encode())).integers(0, numel, 256)


@pytest.mark.parametrize("case", of.FRONTEND_CASES, ids=[c[0] for c in of.FRONTEND_CASES])
def test_tiles_from_u8_golden(golden_dir, case):
    tag, hw, origin, content = case
    g = np.load(os.path.join(golden_dir, "golden_frontend.npz"))
    out = of.tiles_from_u8(of.frontend_image(tag, hw), [origin], content)[0]
    assert out.shape == (3, 1024, 1024) and out.dtype == np.float32
    np.testing.assert_array_equal(out.reshape(-1)[_sample_positions(out.size, tag)], g[f"{tag}.samples"])
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(g[f"{tag}.sha256"])  # bit exact


def test_tiles_from_u8_edges():
    img = of.frontend_image("edge", (300, 200))
    out = of.tiles_from_u8(img, [(0, 0), (250, 150), (300, 0)], (1024, 1024))
    assert np.all(out[0, :, 300:, :] == 0) and np.all(out[0, :, :, 200:] == 0)     # zero padding after normalisation
    assert np.any(out[1, :, :50, :50] != 0) and np.all(out[1, :, 50:, :] == 0)
    assert np.all(out[2] == 0)                                                      # origin outside the image


def test_to_xywh_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_frontend.npz"))
    np.testing.assert_array_equal(of.to_xywh(g["xywh.boxes"]), g["xywh.out"])


def test_merge_detections_properties():
    T, Q = 6, 51
    packed, counts = of.make_tile_detections(T, Q)
    origins = [(0, 0), (0, 896), (896, 0), (896, 896), (1792, 0), (1792, 896)]
    m = of.merge_detections(packed, counts, origins, 0.5)
    n = m["scores"].shape[0]
    assert n == sum(int((packed[t, :counts[t], 4] > np.float32(0.5)).sum()) for t in range(T))
    assert np.all(m["scores"] > np.float32(0.5))
    # tile-major, query order inside a tile
    key = m["src"][:, 0].astype(np.int64) * Q + m["src"][:, 1]
    assert np.all(np.diff(key) > 0)
    t, q = m["src"][5]
    y0, x0 = origins[t]
    np.testing.assert_array_equal(m["boxes"][5], packed[t, q, :4] + np.array([x0, y0, x0, y0], np.float32))
    # cross-tile per-class NMS on top: a detection duplicated in two overlapping tiles collapses to one
    packed2, counts2 = np.zeros((2, 4, 6), np.float32), np.array([1, 1], np.int32)
    packed2[0, 0] = [900, 100, 940, 140, 0.9, 3]
    packed2[1, 0] = [900 - 896, 100, 940 - 896, 140, 0.8, 3]
    m2 = of.merge_detections(packed2, counts2, [(0, 0), (0, 896)], 0.5)
    keep = op.batched_nms(m2["boxes"], m2["scores"], m2["labels"], 0.4)
    np.testing.assert_array_equal(keep, [0])


@pytest.mark.parametrize("case", of.RESIZE_CASES, ids=[c[0] for c in of.RESIZE_CASES])
def test_resize_u8_golden(golden_dir, case):
    """oracle.frontend.resize_u8 vs the golden minted by the reference's RandomResize([768], max_size=768) transform (PIL)."""
    tag, hw, _size, _max = case
    g = np.load(os.path.join(golden_dir, "golden_frontend.npz"))
    oh, ow = (int(v) for v in g[f"{tag}.shape"])
    out = of.resize_u8(of.frontend_image(tag, hw), oh, ow)
    np.testing.assert_array_equal(out.reshape(-1)[_sample_positions(out.size, tag)], g[f"{tag}.samples"])
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(g[f"{tag}.sha256"])  # bit exact


def test_resize_u8_matches_pil_directly():
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    for (H, W, oh, ow) in [(64, 48, 48, 36), (100, 333, 77, 200), (50, 50, 120, 70), (31, 17, 5, 3), (40, 40, 40, 40)]:
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = np.asarray(PIL.fromarray(a).resize((ow, oh), PIL.BILINEAR))
        np.testing.assert_array_equal(of.resize_u8(a, oh, ow), ref)


def test_product_coefficients_equal_oracle():
    """The host side of wm_resize_tiles_u8 (survey.pil_bilinear_coeffs, pure Python doubles) computes Pillow's integers."""
    from wildlifemapper_b200.survey import pil_bilinear_coeffs
    for n_in, n_out in [(1024, 768), (900, 768), (240, 614), (17, 3), (5, 5), (1000, 999)]:
        b, k = pil_bilinear_coeffs(n_in, n_out)
        ob, ok = of.pil_bilinear_coeffs(n_in, n_out)
        np.testing.assert_array_equal(np.array(b, np.int32), ob)
        np.testing.assert_array_equal(np.array(k, np.int32), ok)
