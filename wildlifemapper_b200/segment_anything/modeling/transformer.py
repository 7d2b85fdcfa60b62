"""TwoWayTransformer with the reference's constructor / forward signature / state_dict keys (reference
``modeling/transformer.py``), executed by ``wildlifemapper_b200.engine.DecoderEngine``."""
from typing import Optional, Tuple, Type

import torch
from torch import Tensor, nn

from wildlifemapper_b200.engine import DecoderEngine

from .common import MLPBlock, _MSG, params_version, require_inference, twin_of


class Attention(nn.Module):
    def __init__(self, embedding_dim: int, num_heads: int, downsample_rate: int = 1) -> None:
        super().__init__()
        self.embedding_dim = embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)

    def forward(self, q, k, v):
        raise NotImplementedError(_MSG.format("transformer.Attention"))


class TwoWayAttentionBlock(nn.Module):
    def __init__(self, embedding_dim: int, num_heads: int, mlp_dim: int = 2048,
                 activation: Type[nn.Module] = nn.ReLU, attention_downsample_rate: int = 2,
                 skip_first_layer_pe: bool = False) -> None:
        super().__init__()
        self.self_attn = Attention(embedding_dim, num_heads)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.cross_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.mlp = MLPBlock(embedding_dim, mlp_dim, activation)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe

    def forward(self, queries, keys, query_pe, key_pe):
        raise NotImplementedError(_MSG.format("TwoWayAttentionBlock"))


class TwoWayTransformer(nn.Module):
    def __init__(self, depth: int, embedding_dim: int, num_heads: int, mlp_dim: int,
                 activation: Type[nn.Module] = nn.ReLU, attention_downsample_rate: int = 2) -> None:
        super().__init__()
        if (depth, embedding_dim, num_heads, mlp_dim, attention_downsample_rate) != (2, 256, 8, 2048, 2) \
                or activation is not nn.ReLU:
            raise NotImplementedError("the fused decoder implements the reference factory configuration "
                                      "(depth 2, dim 256, 8 heads, mlp 2048, ReLU, downsample 2), build_sam.py:297-302")
        self.depth = depth
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.mlp_dim = mlp_dim
        self.layers = nn.ModuleList()
        for i in range(depth):
            self.layers.append(TwoWayAttentionBlock(embedding_dim=embedding_dim, num_heads=num_heads, mlp_dim=mlp_dim,
                                                    activation=activation,
                                                    attention_downsample_rate=attention_downsample_rate,
                                                    skip_first_layer_pe=(i == 0)))
        self.final_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)
        self._engine: Optional[DecoderEngine] = None
        self._engine_key = None

    def engine(self) -> DecoderEngine:
        dev = self.norm_final_attn.weight.device
        key = (dev, params_version(self))
        if self._engine is None or self._engine.device != dev:
            self._engine = DecoderEngine(dev)
            self._engine_key = None
        if self._engine_key != key:
            self._engine.prepare_transformer(self.state_dict())
            self._engine_key = key
        return self._engine

    def forward(self, image_embedding: Tensor, image_pe: Tensor, point_embedding: Tensor) -> Tuple[Tensor, Tensor]:
        """image_embedding, image_pe: [B,256,h,w] (64x64); point_embedding [B,Q,256] -> (queries [B,Q,256], keys [B,4096,256])."""
        require_inference(self, image_embedding, image_pe, point_embedding)
        from wildlifemapper_b200.ops import ops
        eng = self.engine()
        B, C, H, W = image_embedding.shape
        assert (C, H, W) == (256, 64, 64), image_embedding.shape
        Q = point_embedding.shape[1]

        def to_tokens(t: Tensor, name: str) -> Tensor:
            rows = twin_of(t)  # token-major twin left by the encoder / get_dense_pe, if still valid
            if rows is not None:
                return rows
            out = eng.ws.get(name, (t.shape[0] * 4096, 256), torch.float32)
            ops.transpose(t.contiguous().float().view(t.shape[0], 256, 4096), out.view(t.shape[0], 4096, 256))
            return out

        feat = to_tokens(image_embedding, "in_feat")
        pe = to_tokens(image_pe, "in_pe")
        tokens = point_embedding.contiguous().float().view(B * Q, 256)
        hs, _, keys = eng.transformer(feat, pe, tokens, B, Q)
        return hs.view(B, Q, 256).clone(), keys.view(B, 4096, 256).clone()
