"""The drop-in HungarianMatcher / SetCriterion (what inference.py:29-89 evaluate() calls for every batch) on the B200, through
the C ABI: matching indices exactly the reference's, losses within 1e-5 of the goldens minted by the reference's own classes
(tests/golden/make_golden.py criterion), cost matrix vs the oracle."""
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wildlifemapper_b200"))

from segment_anything.build_sam import SetCriterion  # noqa: E402  (the drop-in)
from segment_anything.modeling.matcher import HungarianMatcher, build_matcher  # noqa: E402
from segment_anything.utils.misc import reduce_dict  # noqa: E402
from oracle import criterion as oc  # noqa: E402

DEV = "cuda"


def _crit():
    args = types.SimpleNamespace(set_cost_class=1, set_cost_bbox=5, set_cost_giou=2)
    return SetCriterion(7, matcher=build_matcher(args), weight_dict={"loss_ce": 3, "loss_bbox": 5, "loss_giou": 2}, eos_coef=0.1,
                        losses=["labels", "boxes", "cardinality"]).to(DEV).eval()


@pytest.mark.parametrize("case", oc.CRITERION_CASES, ids=[c[0] for c in oc.CRITERION_CASES])
def test_criterion_vs_reference_golden(golden_dir, case):
    tag, B, Q, sizes = case
    g = np.load(os.path.join(golden_dir, "golden_criterion.npz"))
    logits, boxes, targets = oc.make_case(tag, B, Q, sizes)
    out = {"pred_logits": torch.from_numpy(logits).to(DEV), "pred_boxes": torch.from_numpy(boxes).to(DEV)}
    tg = [{"labels": torch.from_numpy(t["labels"]).to(DEV), "boxes": torch.from_numpy(t["boxes"]).to(DEV)} for t in targets]
    crit = _crit()
    with torch.no_grad():
        C = crit.matcher.cost_matrix(out, tg).cpu().numpy()
        idx = crit.matcher(out, tg)
        losses = crit(out, tg)
    if sum(sizes):
        ids = np.concatenate([t["labels"] for t in targets])
        tb = np.concatenate([t["boxes"] for t in targets]).reshape(-1, 4)
        np.testing.assert_allclose(C, oc.match_cost(logits, boxes, ids, tb, 1, 5, 2), rtol=0, atol=3e-6)
    for i, (a, b) in enumerate(idx):  # the reference's assignment, exactly
        assert a.dtype == torch.int64 and not a.is_cuda
        np.testing.assert_array_equal(a.numpy(), g[f"{tag}.idx{i}.src"])
        np.testing.assert_array_equal(b.numpy(), g[f"{tag}.idx{i}.tgt"])
    assert set(losses) == {"loss_ce", "class_error", "loss_bbox", "loss_giou", "cardinality_error"}
    for k, v in losses.items():
        assert v.is_cuda and v.dim() == 0 and v.dtype == torch.float32
        np.testing.assert_allclose(float(v), float(g[f"{tag}.{k}"]), rtol=1e-5, atol=1e-5, err_msg=k)
    # the way evaluate() consumes it (inference.py:52-63)
    red = reduce_dict(losses)
    total = sum(red[k] * crit.weight_dict[k] for k in red if k in crit.weight_dict)
    ref_total = sum(float(g[f"{tag}.{k}"]) * w for k, w in crit.weight_dict.items())
    np.testing.assert_allclose(float(total), ref_total, rtol=1e-5, atol=1e-5)


def test_criterion_on_model_outputs_and_errors():
    """Shaped like one evaluate() step: model outputs -> criterion -> PostProcess; training-style inputs are a loud error."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_model_gpu as tm
    from oracle.weights import make_tiles
    from segment_anything.utils.misc import NestedTensor
    model = tm.build("vit_t", 51)
    tiles = make_tiles(2, seed=2).to(DEV)
    _, _, targets = oc.make_case("c51", 2, 51, (5, 0))
    tg = [{"labels": torch.from_numpy(t["labels"]).to(DEV), "boxes": torch.from_numpy(t["boxes"]).to(DEV)} for t in targets]
    crit = _crit()
    with torch.no_grad():
        out = model(NestedTensor(tiles, None), np.array([[0, 0, 1024, 1024]] * 2))
        losses = crit(out, tg)
    ref_idx = oc.hungarian(oc.match_cost(out["pred_logits"].cpu().numpy(), out["pred_boxes"].cpu().numpy(),
                                         np.concatenate([t["labels"] for t in targets]),
                                         np.concatenate([t["boxes"] for t in targets]).reshape(-1, 4), 1, 5, 2), (5, 0))
    ref = oc.set_criterion(out["pred_logits"].cpu().numpy(), out["pred_boxes"].cpu().numpy(), targets, ref_idx)
    for k, v in ref.items():
        np.testing.assert_allclose(float(losses[k]), v, rtol=1e-5, atol=1e-5, err_msg=k)
    lg = out["pred_logits"].clone().requires_grad_(True)
    with pytest.raises(RuntimeError):
        crit({"pred_logits": lg, "pred_boxes": out["pred_boxes"]}, tg)
    with pytest.raises(RuntimeError):  # CPU tensors: no fallback
        HungarianMatcher(1, 5, 2)({"pred_logits": out["pred_logits"].cpu(), "pred_boxes": out["pred_boxes"].cpu()}, tg)
