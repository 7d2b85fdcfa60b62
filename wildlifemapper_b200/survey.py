"""Either side of the tile-detection path for real survey imagery.

The reference works on pre-cut tile files: PIL image -> ``ToTensor`` -> ``Normalize`` -> zero padding to 1024 x 1024
(``dataloader_coco.py:275-292``, ``utils/misc.py:46-67``), fp32 all the way from the host, and turns detections into
COCO records on the host (``inference.py:149-171``).  Here a whole uint8 survey image (e.g. 5472 x 3648) is uploaded
once (3 bytes per pixel instead of 12), tiles are cut and normalised on the device, and the per-tile detections are
merged into image coordinates with a per-class NMS across overlapping tiles (the cross-tile merge is a north-star
extension: the reference has no equivalent; its oracle is a per-class loop of ``torchvision.ops.nms``).

  plan_tiles(H, W, tile, overlap)      tile origins (host logic, pure Python)
  tiles_from_u8(img, origins, ...)     uint8 HWC image on the device -> fp32 [T,3,1024,1024] normalised, zero padded
  merge_tile_detections(...)           packed per-tile rows -> image-level boxes/scores/labels (+ per-class NMS)
  coco_records(...)                    kept detections -> (xywh+score fp32 [n,5], category int64 [n]) / list of dicts
  SurveyDetector                       the three around a model: image in, COCO records out

``resize_tiles_u8`` reproduces the PIL bilinear resize behind ``RandomResize([768], max_size=768)`` of the reference's
transforms bit-exactly on the device (Pillow's 8-bit fixed-point arithmetic); ``SurveyDetector(resize_to=768)`` cuts
1024 x 1024 tiles, resizes them to 768 x 768 and feeds the reference's "768 x 768 content in a 1024 x 1024 canvas" layout.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import postprocess as pp
from .profiler import ops

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # dataloader_coco.py:281,289
IMAGENET_STD = (0.229, 0.224, 0.225)


def plan_tiles(height: int, width: int, tile: int = 1024, overlap: int = 128) -> List[Tuple[int, int]]:
    """Origins (y0, x0) of ``tile`` x ``tile`` windows covering the image with at least ``overlap`` pixels shared by
    neighbours; the last row / column is shifted back so that it ends at the image border (no partial tiles unless the
    image itself is smaller than a tile)."""
    if height <= 0 or width <= 0:
        raise ValueError("empty image")
    if not 0 <= overlap < tile:
        raise ValueError("overlap must be in [0, tile)")

    def axis(n: int) -> List[int]:
        if n <= tile:
            return [0]
        step = tile - overlap
        pos = list(range(0, n - tile, step))
        pos.append(n - tile)
        return pos

    return [(y, x) for y in axis(height) for x in axis(width)]


def tiles_from_u8(img: torch.Tensor, origins: torch.Tensor, content: Tuple[int, int] = (1024, 1024),
                  mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """img uint8 [H,W,3] on the device, origins int32 [T,2] (y0,x0) on the device -> fp32 [T,3,1024,1024]."""
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
        raise ValueError("img must be uint8 [H,W,3]")
    T = origins.shape[0]
    if out is None:
        out = torch.empty(T, 3, 1024, 1024, device=img.device, dtype=torch.float32)
    ops.tiles_from_u8(img, origins.to(torch.int32).contiguous(), int(content[0]), int(content[1]), list(mean), list(std), out)
    return out


def pil_bilinear_coeffs(in_size: int, out_size: int) -> Tuple[List[Tuple[int, int]], List[List[int]]]:
    """Per output index: (first input index, tap count) and the 22-bit fixed-point taps of Pillow's bilinear resampling
    for 8-bit images (Resample.c: precompute_coeffs with the full-image box + normalize_coeffs_8bpc), evaluated in double
    precision in the same operation order, so the integers are Pillow's."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale  # bilinear: support 1
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, taps = [], []
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)  # (C cast: truncation toward zero)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = []
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            k.append(w)
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        ki = [int(-0.5 + w * (1 << 22)) if w < 0 else int(0.5 + w * (1 << 22)) for w in k]
        bounds.append((xmin, xmax))
        taps.append(ki + [0] * (ksize - xmax))
    return bounds, taps


_COEFF_CACHE: Dict[Tuple[int, int, str], Tuple[torch.Tensor, torch.Tensor]] = {}


def _coeff_tensors(in_size: int, out_size: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    key = (in_size, out_size, str(device))
    if key not in _COEFF_CACHE:
        b, k = pil_bilinear_coeffs(in_size, out_size)
        _COEFF_CACHE[key] = (torch.tensor(b, dtype=torch.int32, device=device), torch.tensor(k, dtype=torch.int32, device=device))
    return _COEFF_CACHE[key]


def resize_tiles_u8(img: torch.Tensor, origins: torch.Tensor, tile: Tuple[int, int], size: Tuple[int, int]) -> torch.Tensor:
    """uint8 [H,W,3] survey image on the device, origins int32 [T,2] (y0,x0) of ``tile`` = (h, w) windows inside it ->
    uint8 [T, size[0], size[1], 3]: every tile resized like ``PIL.Image.resize((w', h'), BILINEAR)`` (what the reference's
    RandomResize does to a tile image), bit-exactly."""
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
        raise ValueError("img must be uint8 [H,W,3]")
    th, tw = int(tile[0]), int(tile[1])
    oh, ow = int(size[0]), int(size[1])
    T = origins.shape[0]
    xb, xk = _coeff_tensors(tw, ow, img.device)
    yb, yk = _coeff_tensors(th, oh, img.device)
    tmp = torch.empty(max(T, 1) * th * ow * 3, device=img.device, dtype=torch.uint8)
    out = torch.empty(T, oh, ow, 3, device=img.device, dtype=torch.uint8)
    ops.resize_tiles_u8(img, origins.to(torch.int32).contiguous(), th, tw, tmp, out, xb, xk, yb, yk)
    return out


def merge_tile_detections(packed: torch.Tensor, counts: torch.Tensor, origins: torch.Tensor, score_thr: float = 0.5,
                          iou_thr: float = 0.4, per_class: bool = True) -> Dict[str, torch.Tensor]:
    """packed fp32 [T,Q,6] / counts int32 [T] (``postprocess_packed`` output, boxes in tile pixels), origins int32 [T,2].
    Returns image-level ``boxes`` [n,4] xyxy, ``scores`` [n], ``labels`` int64 [n], ``src`` int32 [n,2] = (tile, row) of
    every candidate, and ``keep`` int64 [k]: the NMS survivors in score order.  Two device->host reads (candidate and
    kept counts: the result lengths are data dependent)."""
    T, Q, _ = packed.shape
    dev = packed.device
    boxes = torch.empty(max(T * Q, 1), 4, device=dev, dtype=torch.float32)
    scores = torch.empty(max(T * Q, 1), device=dev, dtype=torch.float32)
    labels = torch.empty(max(T * Q, 1), device=dev, dtype=torch.int64)
    src = torch.empty(max(T * Q, 1), 2, device=dev, dtype=torch.int32)
    total = torch.zeros(1, device=dev, dtype=torch.int32)
    tile_n = torch.empty(max(T, 1), device=dev, dtype=torch.int32)
    ops.merge_detections(packed.contiguous(), counts.contiguous(), origins.to(torch.int32).contiguous(), float(score_thr),
                         tile_n, boxes, scores, labels, src, total)
    n = int(total.item())
    boxes, scores, labels, src = boxes[:n], scores[:n], labels[:n], src[:n]
    keep = pp.nms(boxes, scores, iou_thr, labels if per_class else None) if n else torch.empty(0, device=dev, dtype=torch.int64)
    return {"boxes": boxes, "scores": scores, "labels": labels, "src": src, "keep": keep}


def coco_records(boxes: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor, keep: Optional[torch.Tensor] = None):
    """-> (xywh_score fp32 [n,5], category int64 [n]) on the device (``convert_to_xywh``, inference.py:235-237)."""
    n = boxes.shape[0] if keep is None else keep.shape[0]
    dev = boxes.device
    out = torch.empty(max(n, 1), 5, device=dev, dtype=torch.float32)
    cat = torch.empty(max(n, 1), device=dev, dtype=torch.int64)
    if n:
        ops.pack_coco(boxes.contiguous(), scores.contiguous(), labels.contiguous(), keep, n, out, cat)
    return out[:n], cat[:n]


def coco_dicts(image_id, xywh_score: torch.Tensor, category: torch.Tensor) -> List[dict]:
    """The list ``prepare_for_coco_detection`` builds (inference.py:149-171), from one device->host copy."""
    rows, cats = xywh_score.cpu().tolist(), category.cpu().tolist()
    return [{"image_id": image_id, "category_id": c, "bbox": r[:4], "score": r[4]} for r, c in zip(rows, cats)]


class SurveyDetector:
    """uint8 survey image -> image-level detections.  ``model`` is the drop-in ``segment_anything.network.MedSAM`` on a
    B200 in eval mode; tiles run through it in batches of ``batch``."""

    def __init__(self, model, batch: int = 32, tile: int = 1024, overlap: int = 128, conf_thr: float = 0.05,
                 score_thr: float = 0.5, iou_thr: float = 0.4, per_class: bool = True, resize_to: Optional[int] = None):
        if tile > 1024:
            raise ValueError("the encoder takes 1024 x 1024 inputs: tile must be <= 1024")
        if resize_to is not None and not 0 < resize_to <= 1024:
            raise ValueError("resize_to must be in (0, 1024]")
        # resize_to = 768: the reference loader's RandomResize([768], max_size=768) applied to every tile on the device; the
        # model then sees 768 x 768 content in the 1024 x 1024 canvas, boxes still scale by the native tile size (PostProcess
        # multiplies the content-relative boxes by orig_size, inference.py:66)
        self.resize_to = resize_to
        self.model, self.batch, self.tile, self.overlap = model, int(batch), int(tile), int(overlap)
        self.conf_thr, self.score_thr, self.iou_thr, self.per_class = conf_thr, score_thr, iou_thr, per_class
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("SurveyDetector needs the model on a B200 (wildlifemapper_b200 has no CPU path)")

    @torch.no_grad()
    def __call__(self, image_u8: torch.Tensor) -> Dict[str, torch.Tensor]:
        from segment_anything.utils.misc import NestedTensor
        img = image_u8.to(self.device, non_blocking=True)
        H, W, _ = img.shape
        org_host = plan_tiles(H, W, self.tile, self.overlap)
        origins = torch.tensor(org_host, device=self.device, dtype=torch.int32)
        T = origins.shape[0]
        sizes = torch.tensor([[self.tile, self.tile]] * self.batch, device=self.device, dtype=torch.int64)
        buf = torch.empty(self.batch, 3, 1024, 1024, device=self.device, dtype=torch.float32)
        packed_all, counts_all = [], []
        for b0 in range(0, T, self.batch):
            nb = min(self.batch, T - b0)
            if self.resize_to is None:
                tiles = tiles_from_u8(img, origins[b0:b0 + nb], (self.tile, self.tile), out=buf[:nb])
            else:
                r = self.resize_to
                small = resize_tiles_u8(img, origins[b0:b0 + nb], (self.tile, self.tile), (r, r))  # [nb, r, r, 3] uint8
                stacked_org = torch.stack([torch.arange(nb, device=self.device, dtype=torch.int32) * r,
                                           torch.zeros(nb, device=self.device, dtype=torch.int32)], dim=1)
                tiles = tiles_from_u8(small.view(nb * r, r, 3), stacked_org, (r, r), out=buf[:nb])
            out = self.model(NestedTensor(tiles, None), None)
            packed, _l, _q, counts = pp.postprocess_packed(out["pred_logits"], out["pred_boxes"], sizes[:nb], self.conf_thr)
            packed_all.append(packed)
            counts_all.append(counts)
        merged = merge_tile_detections(torch.cat(packed_all), torch.cat(counts_all), origins, self.score_thr, self.iou_thr,
                                       self.per_class)
        merged["origins"] = origins
        return merged
