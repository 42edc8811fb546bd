"""Multi-resolution time-domain feature loss - drop-in mirror of ste_gan/losses/time_domain_loss.py.

`TimeDomainFeatureLoss` / `MultiTimeDomainFeatureLoss` keep the reference's constructor
arguments and methods (`time_domain_loss`, `forward(x_real, x_generated)` - note the
(real, generated) order, time_domain_loss.py:105-107).  The arithmetic - double 9-tap reflect
average, rectified high-pass, framed mean / power features, L1 - is one fused CUDA call
(`stg_td_loss`) for the three resolutions the reference instantiates.
"""
from typing import List, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from ste_gan_b200.layers.average_filter import AverageFilter

_RES = [(20, 8), (51, 13), (80, 16)]      # time_domain_loss.py:88-93


class TimeDomainFeatureLoss(nn.Module):
    """time_domain_loss.py:13-73; only the three (win, shift) pairs of the multi-resolution loss are fused."""

    def __init__(self, num_channels, win_size_samples: int = 21, win_shift_samples: int = 8,
                 apply_padding_windowing: bool = True, average_filter_window_size: int = 9):
        super().__init__()
        self.num_channels = num_channels
        self.average_filter = AverageFilter(num_channels, average_filter_window_size)
        self.win_size_samples = win_size_samples
        self.win_shift_samples = win_shift_samples
        self.apply_padding_windowing = apply_padding_windowing
        self.avg_filter_window_size = average_filter_window_size

    def time_domain_loss(self, x_real: Tensor, x_generated: Tensor):
        from ste_gan_b200.autograd import MultiTdLossFn
        key = (self.win_size_samples, self.win_shift_samples)
        if key not in _RES or not self.apply_padding_windowing or self.avg_filter_window_size != 9:
            raise NotImplementedError(f"TimeDomainFeatureLoss{key}: only {_RES} with padding are on the hot path")
        return MultiTdLossFn.apply(x_real, x_generated)[_RES.index(key)]


class MultiTimeDomainFeatureLoss(nn.Module):
    """time_domain_loss.py:76-107."""

    def __init__(self, num_channels: int):
        super().__init__()
        self.time_domain_losses = nn.ModuleList([
            TimeDomainFeatureLoss(num_channels, win_size_samples=w, win_shift_samples=s) for (w, s) in _RES])

    def time_domain_loss(self, x_real: Tensor, x_generated: Tensor) -> Tuple[Tensor, List[Tensor]]:
        from ste_gan_b200.autograd import MultiTdLossFn
        parts = MultiTdLossFn.apply(x_real, x_generated)
        vals = [parts[0], parts[1], parts[2]]
        return vals[0] + vals[1] + vals[2], vals

    def forward(self, x_real: Tensor, x_generated: Tensor) -> Tensor:
        loss, vals = self.time_domain_loss(x_real, x_generated)
        return loss
