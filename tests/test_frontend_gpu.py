"""Parity of the steps either side of the tile-detection path on the B200, through the
C ABI: bit-exact against oracle/frontend.py (pinned to the reference by tests/test_oracle_frontend.py) and against the
committed golden hashes."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

from wildlifemapper_b200 import survey  # noqa: E402
from oracle import frontend as of  # noqa: E402
from oracle import post as opost  # noqa: E402

DEV = "cuda"


@pytest.mark.parametrize("case", of.FRONTEND_CASES, ids=[c[0] for c in of.FRONTEND_CASES])
def test_tiles_from_u8_golden_bit_exact(golden_dir, case):
    tag, hw, origin, content = case
    g = np.load(os.path.join(golden_dir, "golden_frontend.npz"))
    img = torch.from_numpy(of.frontend_image(tag, hw)).to(DEV)
    out = survey.tiles_from_u8(img, torch.tensor([origin], dtype=torch.int32, device=DEV), content)[0].cpu().numpy()
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(g[f"{tag}.sha256"])


def test_tiles_from_u8_survey_image_bit_exact():
    """Full-size case: a 3648 x 5472 survey image cut into overlapping tiles (incl. a strided view and tiles that hang
    over the border)."""
    H, W = 3648, 5472
    img = of.frontend_image("survey", (H, W))
    origins = survey.plan_tiles(H, W, 1024, 128)
    assert len(origins) == 24 and origins[-1] == (H - 1024, W - 1024)
    pick = [0, 6, 13, 23]
    extra = [(H - 500, W - 300), (H, 0)]  # partly / completely outside: zero padded
    org = [origins[i] for i in pick] + extra
    got = survey.tiles_from_u8(torch.from_numpy(img).to(DEV), torch.tensor(org, dtype=torch.int32, device=DEV)).cpu().numpy()
    np.testing.assert_array_equal(got, of.tiles_from_u8(img, org))
    # a row-strided view of a wider buffer
    wide = torch.zeros(600, 900, 3, dtype=torch.uint8, device=DEV)
    sub = of.frontend_image("sub", (600, 700))
    wide[:, :700] = torch.from_numpy(sub).to(DEV)
    got = survey.tiles_from_u8(wide[:, :700], torch.tensor([[0, 0], [90, 10]], dtype=torch.int32, device=DEV), (768, 768))
    np.testing.assert_array_equal(got.cpu().numpy(), of.tiles_from_u8(sub, [(0, 0), (90, 10)], (768, 768)))


@pytest.mark.parametrize("T,Q", [(6, 51), (28, 900), (1, 1), (300, 51)])
def test_merge_detections_bit_exact(T, Q):
    packed, counts = of.make_tile_detections(T, Q, seed=T + Q)
    rng = np.random.default_rng(1)
    origins = [(int(y), int(x)) for y, x in rng.integers(0, 4000, (T, 2))]
    ref = of.merge_detections(packed, counts, origins, 0.5)
    m = survey.merge_tile_detections(torch.from_numpy(packed).to(DEV), torch.from_numpy(counts).to(DEV),
                                     torch.tensor(origins, dtype=torch.int32, device=DEV), 0.5, 0.4, per_class=True)
    for k in ("boxes", "scores", "labels", "src"):
        np.testing.assert_array_equal(m[k].cpu().numpy(), ref[k])
    np.testing.assert_array_equal(m["keep"].cpu().numpy(), opost.batched_nms(ref["boxes"], ref["scores"], ref["labels"], 0.4))
    xywh, cat = survey.coco_records(m["boxes"], m["scores"], m["labels"], m["keep"])
    k = m["keep"].cpu().numpy()
    np.testing.assert_array_equal(xywh[:, :4].cpu().numpy(), of.to_xywh(ref["boxes"][k]))
    np.testing.assert_array_equal(xywh[:, 4].cpu().numpy(), ref["scores"][k])
    np.testing.assert_array_equal(cat.cpu().numpy(), ref["labels"][k])
    recs = survey.coco_dicts(17, xywh, cat)
    assert len(recs) == k.shape[0] and (not recs or set(recs[0]) == {"image_id", "category_id", "bbox", "score"})


def test_merge_detections_empty_and_duplicates():
    packed = torch.zeros(3, 8, 6, device=DEV)
    counts = torch.zeros(3, dtype=torch.int32, device=DEV)
    m = survey.merge_tile_detections(packed, counts, torch.zeros(3, 2, dtype=torch.int32, device=DEV))
    assert m["boxes"].shape == (0, 4) and m["keep"].shape == (0,)
    xywh, cat = survey.coco_records(m["boxes"], m["scores"], m["labels"], m["keep"])
    assert xywh.shape == (0, 5) and cat.shape == (0,)
    # one animal seen by two overlapping tiles collapses to the higher score; a different class at the same place stays
    p = np.zeros((2, 4, 6), np.float32)
    p[0, 0] = [900, 100, 940, 140, 0.9, 3]
    p[1, 0] = [4, 100, 44, 140, 0.8, 3]
    p[1, 1] = [4, 100, 44, 140, 0.7, 5]
    m = survey.merge_tile_detections(torch.from_numpy(p).to(DEV), torch.tensor([1, 2], dtype=torch.int32, device=DEV),
                                     torch.tensor([[0, 0], [0, 896]], dtype=torch.int32, device=DEV))
    np.testing.assert_array_equal(m["keep"].cpu().numpy(), [0, 2])


def test_survey_detector_end_to_end_matches_tilewise_path():
    """SurveyDetector = front-end + the drop-in model + merge: the same numbers as running the pieces by hand."""
    from test_model_gpu import build  # the drop-in MedSAM with the seeded synthetic weights (also sets sys.path)
    from segment_anything.utils.misc import NestedTensor
    from wildlifemapper_b200 import postprocess as pp
    model = build("vit_t", 51)
    H, W = 1500, 2100
    img = torch.from_numpy(of.frontend_image("e2e", (H, W))).to(DEV)
    det = survey.SurveyDetector(model, batch=4, overlap=128, score_thr=0.0)
    res = det(img)
    origins = survey.plan_tiles(H, W, 1024, 128)
    assert res["origins"].cpu().tolist() == [list(o) for o in origins] and len(origins) == 6
    tiles = torch.from_numpy(of.tiles_from_u8(img.cpu().numpy(), origins)).to(DEV)
    with torch.no_grad():
        out = model(NestedTensor(tiles, None), None)
    packed, _l, _q, counts = pp.postprocess_packed(out["pred_logits"], out["pred_boxes"],
                                                   torch.tensor([[1024, 1024]] * 6, device=DEV), 0.05)
    ref = of.merge_detections(packed.cpu().numpy(), counts.cpu().numpy(), origins, 0.0)
    # batch composition differs (4 + 2 vs 6): kernels are batch-invariant per tile, so this is exact
    np.testing.assert_array_equal(res["src"].cpu().numpy(), ref["src"])
    np.testing.assert_array_equal(res["boxes"].cpu().numpy(), ref["boxes"])
    np.testing.assert_array_equal(res["keep"].cpu().numpy(), opost.batched_nms(ref["boxes"], ref["scores"], ref["labels"], 0.4))


@pytest.mark.parametrize("case", of.RESIZE_CASES, ids=[c[0] for c in of.RESIZE_CASES])
def test_resize_tiles_u8_golden_bit_exact(golden_dir, case):
    """wm_resize_tiles_u8 vs the golden minted through the reference's RandomResize transform (PIL): sha256-exact."""
    tag, hw, _size, _max = case
    g = np.load(os.path.join(golden_dir, "golden_frontend.npz"))
    oh, ow = (int(v) for v in g[f"{tag}.shape"])
    img = torch.from_numpy(of.frontend_image(tag, hw)).to(DEV)
    out = survey.resize_tiles_u8(img, torch.tensor([[0, 0]], dtype=torch.int32, device=DEV), hw, (oh, ow))[0].cpu().numpy()
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(g[f"{tag}.sha256"])


def test_resize_tiles_u8_survey_windows_bit_exact():
    """Several 1024 x 1024 windows of one survey image -> 768 x 768 (the loader's case), against the oracle per tile, and
    chained into tiles_from_u8: the reference's "768 x 768 content in a 1024 x 1024 canvas" tensor, bit-exact."""
    H, W = 2200, 3000
    img = of.frontend_image("survey_rs", (H, W))
    org = [(0, 0), (1176, 1976), (500, 700)]
    small = survey.resize_tiles_u8(torch.from_numpy(img).to(DEV), torch.tensor(org, dtype=torch.int32, device=DEV), (1024, 1024), (768, 768))
    refs = [of.resize_u8(np.ascontiguousarray(img[y:y + 1024, x:x + 1024]), 768, 768) for y, x in org]
    for t in range(len(org)):
        np.testing.assert_array_equal(small[t].cpu().numpy(), refs[t])
    T = len(org)
    stacked = torch.tensor([[t * 768, 0] for t in range(T)], dtype=torch.int32, device=DEV)
    tiles = survey.tiles_from_u8(small.view(T * 768, 768, 3), stacked, (768, 768)).cpu().numpy()
    for t in range(T):
        np.testing.assert_array_equal(tiles[t], of.tiles_from_u8(refs[t], [(0, 0)], (768, 768))[0])
