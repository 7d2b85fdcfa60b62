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

  resize_u8         PIL ``Image.resize(size, BILINEAR)`` on an 8-bit RGB image -- what ``RandomResize([768], max_size=768)``
                    (dataloader_coco.py:275-292 -> utils/augmentation.py:77-107 -> torchvision F.resize) does to a tile --
                    restated from Pillow's src/libImaging/Resample.c (precompute_coeffs, normalize_coeffs_8bpc,
                    ImagingResampleHorizontal_8bpc / Vertical_8bpc; Pillow is a third-party dependency of the reference,
                    12.2.0 in this image): double-precision triangle-filter taps scaled by max(scale, 1), rounded to
                    22-bit integers, horizontal pass into a uint8 intermediate, vertical pass, clip8.

``tests/test_oracle_frontend.py`` pins ``resize_u8`` against PIL itself (and a golden minted through the reference's
``resize`` transform), and pins ``tiles_from_u8`` against a golden minted by running the reference's own
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


# resize cases: tag, tile (H, W), RandomResize size / max_size  (the reference uses ([768], 768) on its tile files)
RESIZE_CASES = (
    ("rs1024", (1024, 1024), 768, 768),   # the loader's case: a 1024 x 1024 tile -> 768 x 768
    ("rs_wide", (600, 900), 768, 768),    # aspect ratio: max_size caps the long side -> (512, 768)
    ("rs_up", (300, 240), 768, 768),      # upscaling (w < h: width -> 614?, decided by get_size_with_aspect_ratio)
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


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs (box = whole image, bilinear filter, support 1) + normalize_coeffs_8bpc.
    -> bounds int32 [out, 2] (xmin, count), kk int32 [out, ksize]."""
    scale = np.float64(in_size) / np.float64(out_size)
    filterscale = max(scale, np.float64(1.0))
    support = np.float64(1.0) * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = np.float64(1.0) / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = np.float64(0.0) + (xx + np.float64(0.5)) * scale
        xmin = max(int(np.trunc(center - support + 0.5)), 0)
        xmax = min(int(np.trunc(center + support + 0.5)), in_size) - xmin
        xs = np.arange(xmax, dtype=np.float64)
        a = np.abs((xs + xmin - center + 0.5) * ss)
        w = np.where(a < 1.0, 1.0 - a, 0.0)
        ww = np.float64(0.0)
        for v in w:  # sequential sum, as in the C loop
            ww = ww + v
        if ww != 0.0:
            w = w / ww
        bounds[xx] = (xmin, xmax)
        kk[xx, :xmax] = np.trunc(np.where(w < 0, -0.5, 0.5) + w * np.float64(1 << 22)).astype(np.int64)
    return bounds, kk


def resize_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H,W,3] -> uint8 [out_h,out_w,3], Pillow's two-pass 8-bit resampling (horizontal first)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, C = img.shape

    def one_pass(src: np.ndarray, n_in: int, n_out: int) -> np.ndarray:  # resamples axis 1 of src [R, n_in, C]
        if n_in == n_out:
            return src
        bounds, kk = pil_bilinear_coeffs(n_in, n_out)
        out = np.empty((src.shape[0], n_out, C), np.uint8)
        s64 = src.astype(np.int64)
        for xx in range(n_out):
            lo, cnt = bounds[xx]
            acc = (1 << 21) + np.tensordot(s64[:, lo:lo + cnt, :], kk[xx, :cnt].astype(np.int64), axes=([1], [0]))
            out[:, xx, :] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
        return out

    tmp = one_pass(img, W, out_w)                                   # [H, out_w, C]
    return one_pass(tmp.transpose(1, 0, 2), H, out_h).transpose(1, 0, 2)  # vertical pass on the transposed view
