"""AverageFilter - drop-in mirror of ste_gan/layers/average_filter.py:10-28.

Reflect-pad by window//2 on both sides followed by AvgPool1d(window, stride 1).  On the hot
path the filter only occurs as the double 9-tap average inside the time-domain loss, where it
is fused into `stg_td_loss`; the stand-alone module is kept for interface parity and runs the
same kernel family through `ste_gan_b200.ops`.
"""
import torch
import torch.nn as nn
from torch import Tensor


class AverageFilter(nn.Module):

    def __init__(self, in_channels: int, window_size: int = 9, pad_signal: bool = True):
        super().__init__()
        if window_size % 2 != 1:
            raise ValueError("window_size must be odd")
        self.in_channels = in_channels
        self.window_size = window_size
        self.padding = window_size // 2
        self.pad_signal = pad_signal

    def forward(self, x: Tensor) -> Tensor:
        """x: [B, C, T] -> [B, C, T] (or [B, C, T - window + 1] without padding)."""
        from .. import ops
        if not x.is_cuda:
            raise RuntimeError("AverageFilter: ste_gan_b200 runs on CUDA only")
        return ops.average_filter(x, self.window_size, self.pad_signal)
