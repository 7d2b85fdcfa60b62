"""Parameter holders with the reference names (``modeling/common.py``).  Inside the fused path their weights are
consumed by ``wildlifemapper_b200.engine``; they are never executed as separate torch modules."""
from typing import Type

import torch
import torch.nn as nn

_MSG = ("{} is a parameter holder in wildlifemapper_b200: its arithmetic is fused into the sm_100a kernel schedule of "
        "the enclosing ImageEncoderViT / MaskDecoder forward; it has no standalone (CPU / eager) path.")


def require_inference(mod: nn.Module, *tensors: torch.Tensor) -> None:
    """The B200 path implements inference only (backward is SURVEY.md section 8f-1, 'next')."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise RuntimeError(
            f"{type(mod).__name__}: inputs require grad, but wildlifemapper_b200 implements the inference path only "
            "(no backward kernels yet). Run under torch.no_grad().")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(f"{type(mod).__name__}: expected CUDA tensors; wildlifemapper_b200 has no CPU fallback")


def params_version(mod: nn.Module):
    """Cheap fingerprint of a module's parameters / buffers: invalidates derived device buffers after
    load_state_dict / .to() / optimizer steps."""
    return tuple((t.data_ptr(), t._version) for t in list(mod.parameters()) + list(mod.buffers()))


def attach_twin(t: torch.Tensor, rows: torch.Tensor, engine=None, gen_attr: str = "") -> None:
    """Hang the token-major twin ``rows`` ([*,4096,C] rows of the NCHW tensor ``t``) on ``t`` so that the next stage can
    skip a transpose round trip.  The twin may live in an engine WORKSPACE that later calls overwrite, and ``t`` may be
    edited in place by the caller, so the attachment carries a validity token: the engine's generation counter for that
    workspace and ``t._version``.  ``twin_of`` returns the twin only while both still match (the reference allows any
    call order: encode A, encode B, decode A must decode A)."""
    gen = getattr(engine, gen_attr) if engine is not None else None
    t._wm_twin = (rows, engine, gen_attr, gen, t._version)


def twin_of(t: torch.Tensor):
    tw = getattr(t, "_wm_twin", None)
    if tw is None:
        return None
    rows, engine, gen_attr, gen, version = tw
    if t._version != version or (engine is not None and getattr(engine, gen_attr) != gen):
        return None
    return rows


class MLPBlock(nn.Module):
    def __init__(self, embedding_dim: int, mlp_dim: int, act: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()

    def forward(self, x):
        raise NotImplementedError(_MSG.format("MLPBlock"))


class LayerNorm2d(nn.Module):
    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps

    def forward(self, x):
        raise NotImplementedError(_MSG.format("LayerNorm2d"))
