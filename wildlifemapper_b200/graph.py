"""CUDA-graph replay of the tile-detection step (no tracing compiler involved: the captured work is exactly the
sequence of hand-written kernels the eager path launches).

The eager path issues ~170 kernels per batch; between them the GPU idles for a few microseconds each (launch latency,
host-side tensor-map encoding).  For a fixed batch shape everything on the path is static -- workspaces, weights and
tensor maps keep their addresses -- so the whole step (fft + encoder + decoder + PostProcess + batched NMS) is captured
once and replayed with a single launch per batch.

    det = GraphedDetector(model, batch=32)          # model: segment_anything.network.MedSAM on a B200, eval mode
    packed, counts, keep_idx, keep_cnt = det(tiles)  # tiles fp32 [32,3,1024,1024] on the same device

Outputs are views of static buffers that the next call overwrites (clone them to keep them).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import postprocess as pp


class GraphedDetector:
    def __init__(self, model, batch: int, target_size=(1024, 1024), conf_thr: float = 0.05, nms_score_thr: float = 0.5,
                 iou_thr: float = 0.4, warmup: int = 2, device: Optional[torch.device] = None, per_class: bool = False,
                 gather: bool = False):
        from segment_anything.utils.misc import NestedTensor  # the drop-in package (same container the callers use)
        self._nested = NestedTensor
        self.model = model
        dev = device if device is not None else next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedDetector needs the model on a B200 (wildlifemapper_b200 has no CPU path)")
        self.device = dev
        self.static_in = torch.zeros(batch, 3, 1024, 1024, device=dev, dtype=torch.float32)
        self.sizes = torch.tensor([list(target_size)] * batch, device=dev, dtype=torch.int64)
        self._thr = (float(conf_thr), float(nms_score_thr), float(iou_thr))
        self._per_class = bool(per_class)
        Q = model.mask_decoder.num_mask_tokens
        self.buffer = pp.DetectionBuffer(batch, Q, dev)  # the four outputs in one flat buffer (one all-gather)
        # gather=True (torch.distributed initialised, world > 1): the NCCL all-gather of the detections is captured in
        # the graph right behind the NMS, so a replay is ONE launch including the exchange
        self.gathered = None
        self._world = 1
        if gather:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self._world = dist.get_world_size()
                self.gathered = torch.zeros(self._world * self.buffer.flat.numel(), device=dev, dtype=torch.float32)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):  # first launches set function attributes / allocate workspaces: not capturable
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.outputs = self._step()

    def _step(self):
        conf, score, iou = self._thr
        out = self.model(self._nested(self.static_in, None), None)
        packed, _labels, _query, counts = pp.postprocess_packed(out["pred_logits"], out["pred_boxes"], self.sizes, conf,
                                                                out=self.buffer)
        keep_idx, keep_cnt = pp.nms_packed(packed, counts, score_thr=score, iou_threshold=iou,
                                           per_class=self._per_class, out=self.buffer)
        if self.gathered is not None:
            from .dist import gather_buffer
            gather_buffer(self.buffer, self.gathered)
        return packed, counts, keep_idx, keep_cnt

    def gathered_outputs(self):
        """(packed [world*B,Q,6], counts, keep_idx, keep_cnt) of ALL ranks after a replay (gather=True)."""
        if self.gathered is None:
            return self.outputs
        return pp.DetectionBuffer.split(self.gathered, self._world, self.buffer.B, self.buffer.Q)

    def replay(self):
        """Run the captured step on whatever ``static_in`` holds."""
        self.graph.replay()
        return self.outputs

    def __call__(self, tiles: torch.Tensor):
        if tuple(tiles.shape) != tuple(self.static_in.shape):
            raise ValueError(f"captured for tiles of shape {tuple(self.static_in.shape)}, got {tuple(tiles.shape)}")
        self.static_in.copy_(tiles, non_blocking=True)
        return self.replay()
