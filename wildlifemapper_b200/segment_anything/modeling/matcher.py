"""HungarianMatcher with the reference's constructor and ``forward(outputs, targets)`` contract (reference
``modeling/matcher.py``).  The matching cost for all queries against all targets of the batch is one sm_100a kernel
(``wm_match_cost``); the linear assignment on the per-image blocks stays scipy's on the host, exactly as in the reference
(``C.cpu()`` then ``linear_sum_assignment``, matcher.py:75-80)."""
import torch
from scipy.optimize import linear_sum_assignment
from torch import nn

from wildlifemapper_b200.ops import ops


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"

    @torch.no_grad()
    def cost_matrix(self, outputs, targets) -> torch.Tensor:
        """fp32 [B, Q, sum T] on the device (matcher.py:58-74)."""
        logits = outputs["pred_logits"]
        if not logits.is_cuda:
            raise RuntimeError("HungarianMatcher: expected CUDA tensors; wildlifemapper_b200 has no CPU fallback")
        bs, nq, c1 = logits.shape
        dev = logits.device
        tgt_ids = torch.cat([v["labels"] for v in targets]).to(device=dev, dtype=torch.int64).contiguous()
        tgt_bbox = torch.cat([v["boxes"] for v in targets]).to(device=dev, dtype=torch.float32).reshape(-1, 4).contiguous()
        cost = torch.empty(bs * nq, tgt_ids.shape[0], device=dev, dtype=torch.float32)
        ops.match_cost(logits.detach().float().reshape(bs * nq, c1).contiguous(),
                       outputs["pred_boxes"].detach().float().reshape(bs * nq, 4).contiguous(), tgt_ids, tgt_bbox,
                       float(self.cost_class), float(self.cost_bbox), float(self.cost_giou), cost)
        return cost.view(bs, nq, -1)

    @torch.no_grad()
    def forward(self, outputs, targets):
        """-> list (one entry per image) of (index_i, index_j) int64 CPU tensors, len = min(num_queries, num_targets)."""
        C = self.cost_matrix(outputs, targets).cpu()
        sizes = [len(v["boxes"]) for v in targets]
        indices = [linear_sum_assignment(c[i]) for i, c in enumerate(C.split(sizes, -1))]
        return [(torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)) for i, j in indices]


def build_matcher(args):
    return HungarianMatcher(cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox, cost_giou=args.set_cost_giou)
