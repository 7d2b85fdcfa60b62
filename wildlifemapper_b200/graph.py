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
                 iou_thr: float = 0.4, warmup: int = 2, device: Optional[torch.device] = None):
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
        packed, _labels, _query, counts = pp.postprocess_packed(out["pred_logits"], out["pred_boxes"], self.sizes, conf)
        keep_idx, keep_cnt = pp.nms_packed(packed, counts, score_thr=score, iou_threshold=iou)
        return packed, counts, keep_idx, keep_cnt

    def replay(self):
        """Run the captured step on whatever ``static_in`` holds."""
        self.graph.replay()
        return self.outputs

    def __call__(self, tiles: torch.Tensor):
        if tuple(tiles.shape) != tuple(self.static_in.shape):
            raise ValueError(f"captured for tiles of shape {tuple(self.static_in.shape)}, got {tuple(tiles.shape)}")
        self.static_in.copy_(tiles, non_blocking=True)
        return self.replay()
