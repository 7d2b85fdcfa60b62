"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the post-process stage.

  postprocess      reference build_sam.py:219-258 (PostProcess.forward) + utils/box_ops.py:9-13
  sigmoid_topk     north-star extension; oracle = sigmoid + STABLE descending sort [:K]
  nms / batched    torchvision.ops.nms semantics as called at visualize_prediction.py:150-154:
                   stable score-descending visit order, fp32 IoU = inter/(a_i+a_j-inter) with no
                   FMA contraction, suppress iff (double)iou > (double)thr;
                   per-class variant = loop of the same over classes (``_batched_nms_vanilla``).

torchvision is a third-party dependency of the reference (pyproject.toml:18-19, uv.lock:813-814
pins 0.23.0; 0.26.0 is what this image has).  ``tests/test_oracle_post.py`` pins this file
against the live ``torchvision.ops.nms`` CPU kernel and the committed golden vectors.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

f32 = np.float32


def softmax_f32(logits: np.ndarray) -> np.ndarray:
    x = logits.astype(f32)
    m = x.max(-1, keepdims=True)
    e = np.exp(x - m, dtype=f32)
    return (e / e.sum(-1, keepdims=True, dtype=f32)).astype(f32)


def select_from_prob(prob: np.ndarray, boxes: np.ndarray, target_sizes: np.ndarray, thr: float = 0.05) -> List[Dict[str, np.ndarray]]:
    """Integer stage of PostProcess given the fp32 class probabilities [B,Q,C+1].

    label = first argmax over the C real classes; keep = score > fp32(thr) (fp32 compare);
    query order preserved; boxes cxcywh->xyxy then * [s0,s1,s0,s1] with s = target_sizes row (h/w swap
    quirk, build_sam.py:252).
    """
    out = []
    thr32 = f32(thr)
    for p, b, ts in zip(prob.astype(f32), boxes.astype(f32), target_sizes):
        s = p[:, :-1].max(-1)
        l = p[:, :-1].argmax(-1).astype(np.int64)  # numpy argmax returns the first maximum
        keep = s > thr32
        cx, cy, w, h = b[keep, 0], b[keep, 1], b[keep, 2], b[keep, 3]
        xyxy = np.stack([cx - f32(0.5) * w, cy - f32(0.5) * h, cx + f32(0.5) * w, cy + f32(0.5) * h], -1).astype(f32)
        img_h, img_w = ts[1], ts[0]
        scale = np.array([img_w, img_h, img_w, img_h]).astype(f32)  # int64 -> fp32 promotion as in torch
        out.append({"scores": s[keep], "labels": l[keep], "boxes": (xyxy * scale).astype(f32),
                    "query": np.nonzero(keep)[0].astype(np.int64)})
    return out


def postprocess(logits: np.ndarray, boxes: np.ndarray, target_sizes: np.ndarray, thr: float = 0.05):
    return select_from_prob(softmax_f32(logits), boxes, target_sizes, thr)


def sigmoid_f32(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32)
    return (f32(1) / (f32(1) + np.exp(-x, dtype=f32))).astype(f32)


def topk_from_prob(prob: np.ndarray, k: int):
    """prob [B, Q*C] fp32 -> (scores [B,k], flat index [B,k]) by stable descending sort."""
    order = np.argsort(-prob, axis=-1, kind="stable")[:, :k]
    return np.take_along_axis(prob, order, -1), order.astype(np.int64)


def sigmoid_topk(logits: np.ndarray, boxes: np.ndarray, k: int, num_classes: int = 7):
    """Deformable-DETR style selection: returns scores, labels, query index, cxcywh boxes."""
    B, Q, _ = logits.shape
    prob = sigmoid_f32(logits[..., :num_classes]).reshape(B, Q * num_classes)
    s, idx = topk_from_prob(prob, k)
    q = idx // num_classes
    return s, (idx % num_classes).astype(np.int64), q, np.take_along_axis(boxes.astype(f32), q[..., None], 1)


def _iou_row(b: np.ndarray, area: np.ndarray, i: int, js: np.ndarray) -> np.ndarray:
    xx1 = np.maximum(b[i, 0], b[js, 0])
    yy1 = np.maximum(b[i, 1], b[js, 1])
    xx2 = np.minimum(b[i, 2], b[js, 2])
    yy2 = np.minimum(b[i, 3], b[js, 3])
    w = np.maximum(f32(0), (xx2 - xx1).astype(f32))
    h = np.maximum(f32(0), (yy2 - yy1).astype(f32))
    inter = (w * h).astype(f32)
    union = ((area[i] + area[js]).astype(f32) - inter).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(f32)


def nms(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """Greedy NMS; returns kept indices (int64) in score-descending (stable) order."""
    b = boxes.astype(f32)
    n = b.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    area = ((b[:, 2] - b[:, 0]).astype(f32) * (b[:, 3] - b[:, 1]).astype(f32)).astype(f32)
    order = np.argsort(-scores.astype(f32), kind="stable")
    bs, ar = b[order], area[order]
    dead = np.zeros(n, bool)
    keep = []
    thr64 = np.float64(thr)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(order[i])
        js = np.arange(i + 1, n)
        iou = _iou_row(bs, ar, i, js)
        dead[js] |= iou.astype(np.float64) > thr64  # NaN compares false, like the C++ kernel
    return np.asarray(keep, np.int64)


def batched_nms(boxes: np.ndarray, scores: np.ndarray, labels: np.ndarray, thr: float) -> np.ndarray:
    """Per-class NMS (loop over classes), result sorted by score descending (stable)."""
    keep_mask = np.zeros(boxes.shape[0], bool)
    for c in np.unique(labels):
        idx = np.nonzero(labels == c)[0]
        keep_mask[idx[nms(boxes[idx], scores[idx], thr)]] = True
    kept = np.nonzero(keep_mask)[0]
    return kept[np.argsort(-scores[kept].astype(f32), kind="stable")].astype(np.int64)


def make_nms_problem(n: int = 10000, seed: int = 3, dup_scores: bool = False):
    """Synthetic dense-herd NMS input."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, 1024, n)
    cy = rng.uniform(0, 1024, n)
    w = np.exp(rng.normal(np.log(32.0), 0.4, n))
    h = np.exp(rng.normal(np.log(32.0), 0.4, n))
    boxes = np.stack([np.clip(cx - w / 2, 0, 1024), np.clip(cy - h / 2, 0, 1024),
                      np.clip(cx + w / 2, 0, 1024), np.clip(cy + h / 2, 0, 1024)], -1).astype(f32)
    scores = ((rng.permutation(n) + 0.5) / n).astype(f32)  # distinct fp32 scores
    if dup_scores:
        scores = (np.floor(scores * 50) / 50).astype(f32)  # heavy ties pin the stable tie-break
    labels = rng.integers(0, 7, n).astype(np.int64)
    return boxes, scores, labels
