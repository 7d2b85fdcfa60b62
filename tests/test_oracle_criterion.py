"""Pin oracle/criterion.py against the golden minted from the reference's own HungarianMatcher and SetCriterion
(tests/golden/make_golden.py criterion).  CPU only."""
import os
import zlib

import numpy as np
import pytest

from oracle import criterion as oc


def _sample_positions(numel, key):  # as tests/golden/make_golden.py
    return np.random.default_rng(zlib.crc32(key.encode())).integers(0, numel, 256)


@pytest.mark.parametrize("case", oc.CRITERION_CASES, ids=[c[0] for c in oc.CRITERION_CASES])
def test_criterion_golden(golden_dir, case):
    tag, B, Q, sizes = case
    g = np.load(os.path.join(golden_dir, "golden_criterion.npz"))
    logits, boxes, targets = oc.make_case(tag, B, Q, sizes)
    ids = np.concatenate([t["labels"] for t in targets])
    tb = np.concatenate([t["boxes"] for t in targets]).reshape(-1, 4)
    C = oc.match_cost(logits, boxes, ids, tb, 1, 5, 2)
    if sum(sizes):
        c = C.reshape(-1)
        np.testing.assert_allclose(c[_sample_positions(c.size, tag + ".cost")], g[f"{tag}.cost.samples"], rtol=0, atol=2e-6)
    idx = oc.hungarian(C, sizes)
    for i, (a, b) in enumerate(idx):  # integer stage: the reference's assignment, exactly
        np.testing.assert_array_equal(a, g[f"{tag}.idx{i}.src"])
        np.testing.assert_array_equal(b, g[f"{tag}.idx{i}.tgt"])
        assert len(a) == min(Q, sizes[i])
    losses = oc.set_criterion(logits, boxes, targets, idx)
    for k, v in losses.items():
        np.testing.assert_allclose(v, float(g[f"{tag}.{k}"]), rtol=2e-6, atol=2e-6, err_msg=k)
