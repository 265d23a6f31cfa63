"""Quantize transform with the reference's constructor (blvm/data/transforms.py:216-260); the bin search runs in a CUDA
kernel and is bit-exact with torch.bucketize(right=False)."""
from typing import Optional

import torch

from . import ops

__all__ = ["Quantize"]


class Quantize(torch.nn.Module):
    def __init__(self, low: float = -1.0, high: float = 1.0, bits: int = 8, bins: Optional[int] = None,
                 force_out_int64: bool = True, rescale: bool = False):
        super().__init__()
        assert (bits is None) != (bins is None), "Must set one and only one of `bits` and `bins`"
        self.low, self.high = low, high
        self.bits = bins // 8 if bits is None else bits
        self.bins = 2 ** bits if bins is None else bins
        # the table is built on the host by torch.linspace exactly like transforms.py:249 (bit-identical boundaries)
        self.register_buffer("boundaries", torch.linspace(start=-1, end=1, steps=self.bins), persistent=False)
        self.out_int32 = (self.bits <= 32) and (not force_out_int64)
        self.rescale = rescale

    def forward(self, x: torch.Tensor):
        if self.boundaries.device != x.device:
            self.boundaries = self.boundaries.to(x.device)
        idx = ops.quantize_indices(x, self.boundaries)
        if self.out_int32:
            idx = idx.to(torch.int32)
        if self.rescale:  # Scale(low, high, min_val=0, max_val=bins-1) of transforms.py:251-252
            return idx.to(torch.float32) / (self.bins - 1) * (self.high - self.low) + self.low
        return idx
