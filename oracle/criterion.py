"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the reference's evaluation-time criterion:

  match_cost        HungarianMatcher.forward's cost matrix (modeling/matcher.py:58-74): softmax probabilities of the target
                    labels, L1 distance of the cxcywh boxes (torch.cdist p=1), generalized IoU of the xyxy boxes
                    (utils/box_ops.py:24-62), weighted sum in the reference's order
  hungarian         the per-image ``scipy.optimize.linear_sum_assignment`` on the diagonal blocks (matcher.py:76-80)
  set_criterion     SetCriterion.forward (build_sam.py:95-210) for losses = labels, boxes, cardinality: weighted cross
                    entropy (empty_weight = 1, ..., 1, eos_coef), class_error = 100 - top-1 accuracy of the matched
                    queries (utils/misc.py:87-102), cardinality error, L1 and GIoU box losses over num_boxes

``tests/test_oracle_criterion.py`` pins it against goldens minted by running the reference's own HungarianMatcher and
SetCriterion (tests/golden/make_golden.py criterion).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

f32 = np.float32


def _xyxy(b: np.ndarray) -> np.ndarray:
    cx, cy, w, h = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([cx - f32(0.5) * w, cy - f32(0.5) * h, cx + f32(0.5) * w, cy + f32(0.5) * h], -1).astype(f32)


def giou_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a [N,4], b [M,4] xyxy fp32 -> [N,M] (box_ops.generalized_box_iou)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt, 0, None).astype(f32)
    inter = wh[..., 0] * wh[..., 1]
    union = (area_a[:, None] + area_b[None, :] - inter).astype(f32)
    iou = (inter / union).astype(f32)
    lt2 = np.minimum(a[:, None, :2], b[None, :, :2])
    rb2 = np.maximum(a[:, None, 2:], b[None, :, 2:])
    wh2 = np.clip(rb2 - lt2, 0, None).astype(f32)
    area = wh2[..., 0] * wh2[..., 1]
    return (iou - (area - union) / area).astype(f32)


def softmax_f32(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32)
    e = np.exp(x - x.max(-1, keepdims=True)).astype(f32)
    return (e / e.sum(-1, keepdims=True, dtype=f32)).astype(f32)


def match_cost(logits: np.ndarray, boxes: np.ndarray, tgt_ids: np.ndarray, tgt_boxes: np.ndarray, w_class: float, w_bbox: float,
               w_giou: float) -> np.ndarray:
    """logits [B,Q,C1], boxes [B,Q,4] cxcywh, targets of the whole batch concatenated -> fp32 [B,Q,T]."""
    B, Q, C1 = logits.shape
    prob = softmax_f32(logits.reshape(B * Q, C1))
    ob = boxes.reshape(B * Q, 4).astype(f32)
    tb = tgt_boxes.reshape(-1, 4).astype(f32)
    cost_class = -prob[:, tgt_ids]
    cost_bbox = np.abs(ob[:, None, :] - tb[None, :, :]).astype(f32).sum(-1, dtype=f32)
    cost_giou = -giou_matrix(_xyxy(ob), _xyxy(tb))
    C = f32(w_bbox) * cost_bbox + f32(w_class) * cost_class + f32(w_giou) * cost_giou
    return C.astype(f32).reshape(B, Q, -1)


def hungarian(C: np.ndarray, sizes: Sequence[int]) -> List[Tuple[np.ndarray, np.ndarray]]:
    from scipy.optimize import linear_sum_assignment
    out, t0 = [], 0
    for i, n in enumerate(sizes):
        r, c = linear_sum_assignment(C[i, :, t0:t0 + n])
        out.append((r.astype(np.int64), c.astype(np.int64)))
        t0 += n
    return out


def set_criterion(logits: np.ndarray, boxes: np.ndarray, targets: List[Dict[str, np.ndarray]], indices, num_classes: int = 7,
                  eos_coef: float = 0.1, world_size: int = 1) -> Dict[str, float]:
    B, Q, C1 = logits.shape
    x = logits.astype(f32)
    tcls = np.full((B, Q), num_classes, np.int64)
    m_logits, m_lab, m_src, m_tgt = [], [], [], []
    for i, (src, J) in enumerate(indices):
        tcls[i, src] = targets[i]["labels"][J]
        m_logits.append(x[i, src])
        m_lab.append(targets[i]["labels"][J])
        m_src.append(boxes[i, src].astype(f32))
        m_tgt.append(targets[i]["boxes"].reshape(-1, 4)[J].astype(f32))
    m_logits, m_lab = np.concatenate(m_logits), np.concatenate(m_lab)
    m_src, m_tgt = np.concatenate(m_src).reshape(-1, 4), np.concatenate(m_tgt).reshape(-1, 4)
    w = np.ones(num_classes + 1, f32)
    w[-1] = f32(eos_coef)
    mx = x.max(-1, keepdims=True)
    logp = (x - mx) - np.log(np.exp(x - mx).sum(-1, keepdims=True, dtype=f32)).astype(f32)
    nll = -np.take_along_axis(logp, tcls[..., None], -1)[..., 0]
    wt = w[tcls]
    loss_ce = float((wt * nll).sum(dtype=np.float64) / wt.sum(dtype=np.float64))
    if m_lab.size:
        acc = float((m_logits[:, :-1].argmax(-1) == m_lab).sum()) * (100.0 / m_lab.size)
    else:
        acc = 0.0
    card_pred = (x.argmax(-1) != C1 - 1).sum(1).astype(f32)
    tgt_len = np.array([len(t["labels"]) for t in targets], f32)
    num_boxes = max(float(sum(len(t["labels"]) for t in targets)) / world_size, 1.0)
    l1 = np.abs(m_src - m_tgt).sum(dtype=np.float64)
    g = np.diag(giou_matrix(_xyxy(m_src), _xyxy(m_tgt))) if m_lab.size else np.zeros((0,), f32)
    return {"loss_ce": loss_ce, "class_error": 100.0 - acc, "cardinality_error": float(np.abs(card_pred - tgt_len).mean()),
            "loss_bbox": float(l1 / num_boxes), "loss_giou": float((1.0 - g.astype(np.float64)).sum() / num_boxes)}


# seeded cases shared by tests/golden/make_golden.py and the tests: tag, B, Q, targets per image
CRITERION_CASES = (
    ("c51", 4, 51, (7, 0, 23, 1)),       # incl. an image without targets
    ("c900", 2, 900, (60, 35)),          # dense herd
    ("cmore", 2, 5, (9, 5)),             # more targets than queries: min(Q, T) matches
    ("cnone", 3, 51, (0, 0, 0)),         # no target anywhere: class_error = 100, num_boxes clamps to 1
)


def make_case(tag: str, B: int, Q: int, sizes: Sequence[int]):
    """Seeded model outputs + targets: logits [B,Q,8], boxes cxcywh in (0,1) with positive extent, labels 1..6."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(tag.encode()))
    logits = (rng.standard_normal((B, Q, 8)) * 2.0).astype(f32)
    cxy = rng.uniform(0.1, 0.9, (B, Q, 2))
    wh = rng.uniform(0.02, 0.2, (B, Q, 2))
    boxes = np.concatenate([cxy, wh], -1).astype(f32)
    targets = []
    for n in sizes:
        t_cxy = rng.uniform(0.1, 0.9, (n, 2))
        t_wh = rng.uniform(0.02, 0.2, (n, 2))
        targets.append({"labels": rng.integers(1, 7, n).astype(np.int64), "boxes": np.concatenate([t_cxy, t_wh], -1).astype(f32)})
    return logits, boxes, targets
