"""Stripped PromptEncoder (reference ``modeling/pos_encoder.py``): only ``pe_layer`` and ``get_dense_pe`` exist.
The dense positional encoding is a constant of the model (it depends only on the [2,128] gaussian buffer), so it
is evaluated once per buffer version and cached; it is not part of the per-tile kernel path."""
from typing import Any, Optional, Tuple, Type

import torch
from torch import nn

from wildlifemapper_b200.engine import DecoderEngine

from .common import attach_twin


class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats: int = 64, scale: Optional[float] = None) -> None:
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))
        self._cache = None

    def dense_tokens(self) -> torch.Tensor:
        """[4096, 256] token-major dense PE for the 64x64 grid."""
        g = self.positional_encoding_gaussian_matrix
        key = (g.data_ptr(), g._version, g.device)
        if self._cache is None or self._cache[0] != key:
            self._cache = (key, DecoderEngine.dense_pe_tokens(g))
        return self._cache[1]

    def forward(self, size: Tuple[int, int]) -> torch.Tensor:
        if tuple(size) != (64, 64):
            raise NotImplementedError("dense PE is specialised to the 64x64 embedding grid")
        tok = self.dense_tokens()
        chw = tok.t().contiguous().view(tok.shape[1], 64, 64)
        attach_twin(chw, tok)
        return chw


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim: int, image_embedding_size: Tuple[int, int], input_image_size: Tuple[int, int],
                 mask_in_chans: int, activation: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self._pe = None

    def get_dense_pe(self) -> torch.Tensor:
        """1 x embed_dim x 64 x 64 (reference pos_encoder.py:24-33); cached, carries its token-major copy."""
        tok = self.pe_layer.dense_tokens()
        if self._pe is None or self._pe[0] is not tok:
            pe = tok.t().contiguous().view(1, tok.shape[1], 64, 64)
            attach_twin(pe, tok)
            self._pe = (tok, pe)
        return self._pe[1]
