"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement of the steps either side of the tile-detection path:

  tiles_from_u8     torchvision ``to_tensor`` + ``normalize`` as composed by the reference's transforms
                    (dataloader_coco.py:286-292; utils/augmentation.py:229-231 ToTensor, :243-249 Normalize):
                    ``(u8.float() / 255 - mean) / std`` in fp32, then the zero padding into a [3,1024,1024] canvas of
                    ``nested_tensor_from_tensor_list`` (utils/misc.py:46-67; padding is zero AFTER normalisation,
                    content is clipped at 1024).  Tiles are cut from one HWC uint8 image at (y0, x0) origins.
  merge_detections  score filter ``scores > thr`` (visualize_prediction.py:150) over the PostProcess rows of every tile,
                    tile-major / query order, boxes moved by the tile origin (fp32 add); the cross-tile NMS on top
                    of it is ``oracle.post.batched_nms`` (per-class loop of torchvision.ops.nms -- north-star
                    extension).
  to_xywh           ``convert_to_xywh`` (inference.py:235-237): (xmin, ymin, xmax - xmin, ymax - ymin).

``tests/test_oracle_frontend.py`` pins ``tiles_from_u8`` against a golden minted by running the reference's own
transform classes and ``nested_tensor_from_tensor_list`` (tests/golden/make_golden.py), and ``to_xywh`` against the torch
expression of the reference.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np

f32 = np.float32
CANVAS = 1024
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def tiles_from_u8(img: np.ndarray, origins: Sequence[Tuple[int, int]], content: Tuple[int, int] = (CANVAS, CANVAS),
                  mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD) -> np.ndarray:
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
    H, W, _ = img.shape
    m = np.asarray(mean, dtype=f32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=f32).reshape(3, 1, 1)
    out = np.zeros((len(origins), 3, CANVAS, CANVAS), f32)
    for t, (y0, x0) in enumerate(origins):
        h = max(0, min(content[0], CANVAS, H - y0))
        w = max(0, min(content[1], CANVAS, W - x0))
        if h == 0 or w == 0:
            continue
        crop = img[y0:y0 + h, x0:x0 + w].transpose(2, 0, 1).astype(f32)
        out[t, :, :h, :w] = ((crop / f32(255)) - m) / s
    return out


def merge_detections(packed: np.ndarray, counts: np.ndarray, origins: Sequence[Tuple[int, int]], thr: float) -> Dict[str, np.ndarray]:
    """packed fp32 [T,Q,6] = x0,y0,x1,y1,score,label; counts [T]."""
    boxes, scores, labels, src = [], [], [], []
    thr32 = f32(thr)
    for t, (y0, x0) in enumerate(origins):
        rows = packed[t, : int(counts[t])].astype(f32)
        keep = rows[:, 4] > thr32
        off = np.array([x0, y0, x0, y0]).astype(f32)
        boxes.append((rows[keep, :4] + off).astype(f32))
        scores.append(rows[keep, 4])
        labels.append(rows[keep, 5].astype(np.int64))
        q = np.nonzero(keep)[0]
        src.append(np.stack([np.full_like(q, t), q], -1).astype(np.int32))
    return {"boxes": np.concatenate(boxes).reshape(-1, 4) if boxes else np.zeros((0, 4), f32),
            "scores": np.concatenate(scores) if scores else np.zeros((0,), f32),
            "labels": np.concatenate(labels) if labels else np.zeros((0,), np.int64),
            "src": np.concatenate(src).reshape(-1, 2) if src else np.zeros((0, 2), np.int32)}


def to_xywh(boxes: np.ndarray) -> np.ndarray:
    b = boxes.astype(f32).reshape(-1, 4)
    return np.stack([b[:, 0], b[:, 1], b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], -1).astype(f32)


# seeded inputs shared by tests/golden/make_golden.py and the tests: tag, image (H, W), origin (y0, x0), content (h, w)
FRONTEND_CASES = (
    ("small", (37, 53), (0, 0), (1024, 1024)),
    ("clip", (1100, 1200), (40, 100), (1024, 1024)),
    ("c768", (900, 800), (100, 20), (768, 768)),
)


def frontend_image(tag: str, hw) -> np.ndarray:
    import zlib
    return np.random.default_rng(zlib.crc32(tag.encode())).integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)


def make_tile_detections(T: int, Q: int, seed: int = 7):
    """Synthetic PostProcess output for T tiles: packed fp32 [T,Q,6], counts int32 [T] (some empty, some full)."""
    rng = np.random.default_rng(seed)
    packed = np.zeros((T, Q, 6), f32)
    xy = rng.uniform(0, 1000, (T, Q, 2)).astype(f32)
    wh = np.exp(rng.normal(np.log(32), 0.4, (T, Q, 2))).astype(f32)
    packed[..., 0:2] = xy
    packed[..., 2:4] = np.minimum(xy + wh, f32(1024))
    packed[..., 4] = rng.uniform(0.05, 1.0, (T, Q)).astype(f32)
    packed[..., 5] = rng.integers(0, 7, (T, Q)).astype(f32)
    counts = rng.integers(0, Q + 1, T).astype(np.int32)
    counts[0] = 0
    counts[-1] = Q
    return packed, counts
