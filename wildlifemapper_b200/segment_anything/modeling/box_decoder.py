"""MaskDecoder (really the box / class decoder) with the reference's constructor, forward signature and
state_dict keys (reference ``modeling/box_decoder.py:15-176``), executed by ``DecoderEngine``."""
from typing import Optional, Type

import torch
from torch import nn

from .common import _MSG, params_version, require_inference, twin_of


class MLP(nn.Module):
    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int,
                 sigmoid_output: bool = False) -> None:
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output

    def forward(self, x):
        raise NotImplementedError(_MSG.format("MLP"))


class MaskDecoder(nn.Module):
    def __init__(self, *, transformer_dim: int, transformer: nn.Module, num_multimask_outputs: int = 3,
                 activation: Type[nn.Module] = nn.GELU, iou_head_depth: int = 3, iou_head_hidden_dim: int = 256,
                 aux_loss=False, embed_dim=256) -> None:
        super().__init__()
        if transformer_dim != 256 or iou_head_hidden_dim != 256 or iou_head_depth != 3:
            raise NotImplementedError("the fused decoder is specialised to transformer_dim 256 / 3-layer heads")
        if aux_loss:
            raise NotImplementedError("aux_loss is a training-only option (out of scope, SURVEY.md section 8f)")
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.aux_loss = aux_loss
        self.num_classes = 6 + 1
        self.iou_token = nn.Embedding(1, transformer_dim)  # unused, kept for the state_dict (box_decoder.py:52)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.class_embed = MLP(transformer_dim, iou_head_hidden_dim, self.num_classes + 1, 3)
        self.bbox_embed = MLP(transformer_dim, iou_head_hidden_dim, 4, 3)
        self._heads_key = None

    def forward(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse_prompt_embeddings=None,
                dense_prompt_embeddings=None, multimask_output: bool = False, hfc_embed=None):
        """-> {'pred_logits': fp32 [B,Q,8], 'pred_boxes': fp32 [B,Q,4]}.  As in the reference, the prompt / hfc
        arguments are accepted and ignored (box_decoder.py:71-107,128-139)."""
        require_inference(self, image_embeddings, image_pe)
        from wildlifemapper_b200.ops import ops
        eng = self.transformer.engine()
        key = params_version(self)
        if self._heads_key != (key, id(eng)):
            eng.prepare_heads(self.state_dict())
            self._heads_key = (key, id(eng))
        B = image_embeddings.shape[0]
        Q = self.num_mask_tokens

        def to_tokens(t: torch.Tensor, name: str) -> torch.Tensor:
            rows = twin_of(t)  # token-major twin left by the encoder / get_dense_pe, if still valid
            if rows is not None:
                return rows
            out = eng.ws.get(name, (t.shape[0] * 4096, 256), torch.float32)
            ops.transpose(t.contiguous().float().view(t.shape[0], 256, 4096), out.view(t.shape[0], 4096, 256))
            return out

        feat = to_tokens(image_embeddings, "in_feat")
        pe = to_tokens(image_pe, "in_pe")  # [4096,256]: broadcast over the batch like repeat_interleave at :139
        _, hsb, _ = eng.transformer(feat, pe, eng.w["tokens"], B, Q)
        logits, boxes = eng.heads(hsb, B, Q)
        return {"pred_logits": logits, "pred_boxes": boxes}
