"""Pin oracle/post.py against (a) golden vectors minted from the reference PostProcess and torchvision NMS
(tests/golden/make_golden.py) and (b) the live torchvision CPU kernel when present.  CPU only."""
import os

import numpy as np
import pytest

from oracle import post as op


def test_postprocess_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_post.npz"))
    for case in range(3):
        logits, boxes, sizes = g[f"c{case}.logits"], g[f"c{case}.boxes"], g[f"c{case}.sizes"]
        res = op.postprocess(logits, boxes, sizes, 0.05)
        for i, r in enumerate(res):
            ref_s, ref_l, ref_b = g[f"c{case}.{i}.scores"], g[f"c{case}.{i}.labels"], g[f"c{case}.{i}.boxes"]
            assert r["labels"].shape == ref_l.shape, (case, i)
            np.testing.assert_array_equal(r["labels"], ref_l)          # integer stage: bit exact
            np.testing.assert_allclose(r["scores"], ref_s, rtol=0, atol=5e-7)  # float stage: 2 ulp of fp32 softmax
            np.testing.assert_allclose(r["boxes"], ref_b, rtol=0, atol=1e-3)
    # the empty-result branch exists in the fixture (case 2, image 0)
    assert g["c2.0.labels"].shape[0] == 0 and g["c2.1.labels"].shape[0] >= 1


@pytest.mark.parametrize("tag,n,dup", [("n2000", 2000, False), ("n2000dup", 2000, True), ("n10000", 10000, False)])
def test_nms_golden(golden_dir, tag, n, dup):
    g = np.load(os.path.join(golden_dir, "golden_nms.npz"))
    b, s, l = op.make_nms_problem(n, seed=3, dup_scores=dup)
    np.testing.assert_array_equal(op.nms(b, s, 0.4), g[f"{tag}.keep"].astype(np.int64))
    np.testing.assert_array_equal(op.batched_nms(b, s, l, 0.4), g[f"{tag}.keep_per_class"].astype(np.int64))


def test_nms_threshold_tie_is_double_compare(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_nms.npz"))
    np.testing.assert_array_equal(op.nms(g["tie.boxes"], g["tie.scores"], 0.4), g["tie.keep"].astype(np.int64))
    # IoU == fp32(2/5) must be suppressed at thr=0.4
    assert 1 not in op.nms(g["tie.boxes"], g["tie.scores"], 0.4)


def test_nms_live_torchvision():
    tv = pytest.importorskip("torchvision")
    import torch
    for seed in range(3):
        b, s, _ = op.make_nms_problem(500, seed=seed, dup_scores=bool(seed % 2))
        ref = tv.ops.nms(torch.from_numpy(b), torch.from_numpy(s), 0.4).numpy()
        np.testing.assert_array_equal(op.nms(b, s, 0.4), ref)


def test_nms_edge_cases():
    assert op.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.4).shape == (0,)
    one = np.array([[0, 0, 1, 1]], np.float32)
    np.testing.assert_array_equal(op.nms(one, np.array([0.3], np.float32), 0.4), [0])
    # degenerate zero-area duplicates: IoU = 0/0 = NaN -> never suppressed
    z = np.array([[5, 5, 5, 5], [5, 5, 5, 5]], np.float32)
    np.testing.assert_array_equal(op.nms(z, np.array([0.9, 0.8], np.float32), 0.4), [0, 1])


def test_sigmoid_topk_stable():
    logits = np.zeros((1, 4, 8), np.float32)
    logits[0, :, :7] = -3
    logits[0, 2, 1] = logits[0, 0, 5] = logits[0, 3, 0] = 1.5   # three exact ties
    boxes = np.random.default_rng(0).random((1, 4, 4)).astype(np.float32)
    s, l, q, bx = op.sigmoid_topk(logits, boxes, 5)
    assert list(q[0][:3]) == [0, 2, 3] and list(l[0][:3]) == [5, 1, 0]   # lower flat index first
    np.testing.assert_array_equal(bx[0, 0], boxes[0, 0])
