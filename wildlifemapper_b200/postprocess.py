"""Front-ends of the post-process kernels.

  postprocess_packed   PostProcess.forward semantics (reference build_sam.py:219-258) -> packed [B,Q,6] + counts
  sigmoid_topk         north-star extension (Deformable-DETR style selection)
  nms / batched_nms    torchvision.ops.nms semantics (reference call site visualize_prediction.py:150-154)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .profiler import ops


class DetectionBuffer:
    """The four fixed-shape detection outputs of a batch (packed rows, valid counts, NMS keep lists, keep counts) carved
    out of ONE flat 4-byte-element buffer, so that a rank ships them with a single NCCL all-gather
    (wildlifemapper_b200.dist.gather_buffer)."""

    def __init__(self, B: int, Q: int, device):
        self.B, self.Q = B, Q
        n_p, n_k = B * Q * 6, B * Q
        self.flat = torch.zeros(n_p + B + n_k + B, device=device, dtype=torch.float32)
        self.packed = self.flat[:n_p].view(B, Q, 6)
        self.counts = self.flat[n_p:n_p + B].view(torch.int32)
        self.keep_idx = self.flat[n_p + B:n_p + B + n_k].view(torch.int32).view(B, Q)
        self.keep_cnt = self.flat[n_p + B + n_k:].view(torch.int32)

    @staticmethod
    def split(flat: torch.Tensor, world: int, B: int, Q: int):
        """[world * n] gathered flat buffers -> (packed [world*B,Q,6], counts [world*B], keep_idx [world*B,Q], keep_cnt
        [world*B]), rank-major = global tile order for contiguous shards."""
        n_p, n_k = B * Q * 6, B * Q
        f = flat.view(world, -1)
        return (f[:, :n_p].reshape(world * B, Q, 6), f[:, n_p:n_p + B].reshape(-1).view(torch.int32),
                f[:, n_p + B:n_p + B + n_k].reshape(world * B, Q).view(torch.int32),
                f[:, n_p + B + n_k:].reshape(-1).view(torch.int32))


def postprocess_packed(logits: torch.Tensor, boxes: torch.Tensor, target_sizes: torch.Tensor, thr: float = 0.05,
                       from_prob: bool = False, out: Optional[DetectionBuffer] = None):
    """Returns (packed fp32 [B,Q,6] = x1,y1,x2,y2,score,label; labels int64 [B,Q]; query int32 [B,Q]; counts int32 [B]).
    Rows [0, counts[b]) of image b are valid, in query order.  ``out``: write packed / counts into that buffer."""
    B, Q, _ = logits.shape
    dev = logits.device
    packed = out.packed if out is not None else torch.empty(B, Q, 6, device=dev, dtype=torch.float32)
    labels = torch.empty(B, Q, device=dev, dtype=torch.int64)
    query = torch.empty(B, Q, device=dev, dtype=torch.int32)
    counts = out.counts if out is not None else torch.empty(B, device=dev, dtype=torch.int32)
    ops.postprocess(logits.contiguous().float(), boxes.contiguous().float(),
                    target_sizes.to(device=dev, dtype=torch.int64).contiguous(), float(thr), int(from_prob), packed,
                    query, labels, counts)
    return packed, labels, query, counts


def sigmoid_topk(logits: torch.Tensor, boxes: torch.Tensor, k: int, num_classes: int = 7,
                 prob: Optional[torch.Tensor] = None):
    """scores fp32 [B,k], labels int32 [B,k], query int32 [B,k], boxes fp32 [B,k,4] (cxcywh).
    ``prob`` (fp32 [B,Q*num_classes]) bypasses the in-kernel sigmoid (integer-stage parity tests)."""
    B, Q, _ = logits.shape
    dev = logits.device
    n = Q * num_classes
    prob_ws = torch.empty(B, n, device=dev, dtype=torch.float32) if prob is None else prob.contiguous().clone()
    order_ws = torch.empty(B, n, device=dev, dtype=torch.int32)
    scores = torch.empty(B, k, device=dev, dtype=torch.float32)
    labels = torch.empty(B, k, device=dev, dtype=torch.int32)
    query = torch.empty(B, k, device=dev, dtype=torch.int32)
    out_boxes = torch.empty(B, k, 4, device=dev, dtype=torch.float32)
    ops.sigmoid_topk(logits.contiguous().float(), boxes.contiguous().float(), prob_ws, order_ws, scores, labels, query,
                     out_boxes, num_classes, k, int(prob is not None))
    return scores, labels, query, out_boxes


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float,
        labels: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Kept indices (int64, score-descending, stable).  ``labels`` given -> per-class NMS.
    One device->host read of the kept count (the result length is data dependent, as in torchvision)."""
    n = boxes.shape[0]
    dev = boxes.device
    keep = torch.empty(max(n, 1), device=dev, dtype=torch.int64)
    num = torch.zeros(1, device=dev, dtype=torch.int32)
    order_ws = torch.empty(max(n, 1), device=dev, dtype=torch.int32)
    mask_ws = torch.empty(max(n * ((n + 63) // 64), 1), device=dev, dtype=torch.int64)
    ops.nms(boxes.contiguous().float().view(-1, 4), scores.contiguous().float(),
            None if labels is None else labels.to(torch.int64).contiguous(), float(iou_threshold), order_ws, mask_ws, keep,
            num)
    return keep[: int(num.item())]


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    return nms(boxes, scores, iou_threshold, labels)


def nms_packed(packed: torch.Tensor, counts: torch.Tensor, score_thr: float = 0.5, iou_threshold: float = 0.4,
               per_class: bool = False, out: Optional[DetectionBuffer] = None):
    """Batched NMS over PostProcess output, entirely on the device (no host sync).
    Returns (keep_idx int32 [B,Q] row indices in score order, keep_cnt int32 [B])."""
    B, Q, _ = packed.shape
    keep_idx = out.keep_idx if out is not None else torch.empty(B, Q, device=packed.device, dtype=torch.int32)
    keep_cnt = out.keep_cnt if out is not None else torch.empty(B, device=packed.device, dtype=torch.int32)
    ops.nms_batched(packed, counts, float(score_thr), float(iou_threshold), int(per_class), keep_idx, keep_cnt)
    return keep_idx, keep_cnt
