"""Multi-GPU plumbing: tiles are independent units (no cross-sample op anywhere on the path), so a global batch
is split contiguously across ranks with a full weight replica per rank and NO data-path collective; only the
packed detections are gathered (reference analogue: the pickled all_gather in
utils/misc.py:180-220 used by inference.py:240-259).  Fixed-shape tensors -> one NCCL all-gather over NVLink."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (n % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(packed: torch.Tensor, counts: torch.Tensor, keep_idx: Optional[torch.Tensor] = None,
                      keep_cnt: Optional[torch.Tensor] = None, group=None):
    """All-gather the fixed-shape detection buffers of every rank (equal local batch per rank).

    packed fp32 [B,Q,6], counts int32 [B] (+ optional NMS keep lists int32 [B,Q] / [B]).  Returns tensors with
    a leading world*B dimension, rank-major (= global tile order for contiguous shards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return packed, counts, keep_idx, keep_cnt
    world = dist.get_world_size(group)

    def ag(t: Optional[torch.Tensor]):
        if t is None:
            return None
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
        if t.is_cuda:
            dist.all_gather_into_tensor(out, t.contiguous(), group=group)  # one NCCL all-gather
        else:  # gloo (CPU tests of the host logic)
            dist.all_gather(list(out.chunk(world, dim=0)), t.contiguous(), group=group)
        return out

    return ag(packed), ag(counts), ag(keep_idx), ag(keep_cnt)


def gather_buffer(buf, out_flat: Optional[torch.Tensor] = None, group=None):
    """ONE all-gather of a ``postprocess.DetectionBuffer`` (all four detection tensors travel in one flat buffer).
    Capturable in a CUDA graph (fixed addresses when ``out_flat`` is given).  Returns the gathered flat buffer
    [world * n]; ``DetectionBuffer.split`` turns it into the four tensors."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return buf.flat
    world = dist.get_world_size(group)
    if out_flat is None:
        out_flat = torch.empty(world * buf.flat.numel(), device=buf.flat.device, dtype=buf.flat.dtype)
    if buf.flat.is_cuda:
        dist.all_gather_into_tensor(out_flat, buf.flat, group=group)
    else:
        dist.all_gather(list(out_flat.chunk(world, dim=0)), buf.flat, group=group)
    return out_flat
