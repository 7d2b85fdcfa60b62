"""``Sam`` container (reference ``modeling/sam.py:18-50``): the factory builds it and callers read its three
sub-modules (train.py:193-197).  Its upstream ``forward`` is dead code in the reference (it calls the modified
encoder without ``x_hfc`` and would raise) and is not part of the hot path."""
from typing import Any, List

import torch
from torch import nn


class Sam(nn.Module):
    mask_threshold: float = 0.0
    image_format: str = "RGB"

    def __init__(self, image_encoder, prompt_encoder, mask_decoder,
                 pixel_mean: List[float] = [123.675, 116.28, 103.53],
                 pixel_std: List[float] = [58.395, 57.12, 57.375]) -> None:
        super().__init__()
        self.image_encoder = image_encoder
        self.prompt_encoder = prompt_encoder
        self.mask_decoder = mask_decoder
        self.register_buffer("pixel_mean", torch.Tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.Tensor(pixel_std).view(-1, 1, 1), False)

    @property
    def device(self) -> Any:
        return self.pixel_mean.device

    def forward(self, *args, **kwargs):
        raise NotImplementedError("Sam.forward is dead code in the reference (sam.py:99 omits x_hfc); use "
                                  "segment_anything.network.MedSAM")
