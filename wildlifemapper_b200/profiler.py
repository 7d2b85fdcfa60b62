"""CUDA-event kernel timer used by bench.py: brackets every ``torch.ops.wm_b200.*`` launch made by the engines
with events on the launching stream and attributes algorithmic work (FLOPs for tensor-core kernels, bytes for
bandwidth kernels) to each kernel family.  Inactive (zero overhead beyond one attribute
lookup) unless ``start()`` has been called."""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional

import torch

from .ops import ops as _raw_ops

_ACTIVE: Optional["KernelTimer"] = None
_CREDIT = 1.0  # fraction of a launch's executed work that is ALGORITHMIC (see credit())


class credit:
    """Context manager: launches inside it are credited `fraction` of their executed FLOPs / bytes.  Used where the engine
    executes more than the algorithm asks for (the hi + lo split products of the low-pass GEMMs: 3 products per stage for
    one algorithmic product), so that roofline fractions are not inflated by it."""

    def __init__(self, fraction: float):
        self.fraction = float(fraction)

    def __enter__(self):
        global _CREDIT
        self.prev, _CREDIT = _CREDIT, self.fraction
        return self

    def __exit__(self, *exc):
        global _CREDIT
        _CREDIT = self.prev
        return False


def _nbytes(*ts) -> int:
    return sum(t.numel() * t.element_size() for t in ts if isinstance(t, torch.Tensor))


def _work(name: str, a) -> tuple:
    """-> (kind, amount): kind 'flop' or 'byte' (algorithmic)."""
    if name == "gemm":
        (M, K), N = a[0].shape, a[1].shape[0]
        return "flop", 2.0 * M * N * K
    if name == "conv3x3":
        B, _, _, C = a[0].shape
        return "flop", 2.0 * B * 4096 * a[1].shape[0] * 9 * C
    if name == "attn_flash":
        B, H, Tq, Tk, hd = a[8], a[9], a[10], a[11], a[12]
        f = 4.0 * B * H * Tq * Tk * hd
        if a[6] is not None:
            f += 4.0 * B * H * Tq * 64 * hd  # q.Rh / q.Rw table products
        return "flop", f
    if name == "attn_window":
        out, H = a[2], a[3]
        D = out.shape[-1]
        B = out.numel() // (4096 * D)
        return "flop", B * (4.0 * 4096 * 196 * D + 4.0 * 4096 * 14 * D)
    if name == "attn_small":
        B, H, Tq, Tk, hd = a[4], a[5], a[6], a[7], a[8]
        return "flop", 4.0 * B * H * Tq * Tk * hd
    return "byte", float(_nbytes(*a))


class KernelTimer:
    def __init__(self) -> None:
        self.records: List[tuple] = []

    def summary(self) -> Dict[str, dict]:
        torch.cuda.synchronize()
        agg: Dict[str, dict] = defaultdict(lambda: {"launches": 0, "ms": 0.0, "flop": 0.0, "byte": 0.0})
        for name, s, e, kind, amount in self.records:
            r = agg[name]
            r["launches"] += 1
            r["ms"] += s.elapsed_time(e)
            r[kind] += amount
        return dict(agg)


def start() -> KernelTimer:
    global _ACTIVE
    _ACTIVE = KernelTimer()
    return _ACTIVE


def stop() -> None:
    global _ACTIVE
    _ACTIVE = None


class _TimedOps:
    def __getattr__(self, name: str):
        op = getattr(_raw_ops, name)

        def call(*args):
            t = _ACTIVE
            if t is None:
                return op(*args)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            op(*args)
            e.record()
            kind, amount = _work(name, args)
            t.records.append((name, s, e, kind, amount * _CREDIT))

        setattr(self, name, call)  # cache the wrapper
        return call


ops = _TimedOps()
