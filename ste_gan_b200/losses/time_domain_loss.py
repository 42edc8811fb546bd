"""Multi-resolution time-domain feature loss - drop-in mirror of ste_gan/losses/time_domain_loss.py.

`TimeDomainFeatureLoss` / `MultiTimeDomainFeatureLoss` keep the reference's constructor
arguments and methods (`window_signal`, `frame_means`, `frame_power`, `double_average`,
`calculate_time_domain_features`, `time_domain_loss`, `forward(x_real, x_generated)` - note the
(real, generated) order, time_domain_loss.py:105-107).  The arithmetic - double reflect
average, rectified high-pass, framed mean / power features, L1 - is one fused CUDA call
(`stg_td_loss_ex`; the three resolutions the reference instantiates share one call).
"""
from typing import List, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from ste_gan_b200.layers.average_filter import AverageFilter

_RES = [(20, 8), (51, 13), (80, 16)]      # time_domain_loss.py:88-93


class TimeDomainFeatureLoss(nn.Module):
    """time_domain_loss.py:13-73 - every constructor setting is supported (any window / shift, with or without the
    reflect-padded windowing, any odd average-filter window) through `stg_td_loss_ex`; the helper methods of the
    reference class are kept and run the same CUDA kernels (forward only, like the reference's use of them)."""

    def __init__(self, num_channels, win_size_samples: int = 21, win_shift_samples: int = 8,
                 apply_padding_windowing: bool = True, average_filter_window_size: int = 9):
        super().__init__()
        self.num_channels = num_channels
        self.average_filter = AverageFilter(num_channels, average_filter_window_size)
        self.win_size_samples = win_size_samples
        self.win_shift_samples = win_shift_samples
        self.apply_padding_windowing = apply_padding_windowing
        self.avg_filter_window_size = average_filter_window_size

    @staticmethod
    def _cuda(x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("TimeDomainFeatureLoss: ste_gan_b200 runs on CUDA only")
        return x.detach()

    def window_signal(self, _x: Tensor) -> Tensor:
        """[B,T,C] -> [B,F,C,win] frames (time_domain_loss.py:35-41)."""
        from ste_gan_b200 import ops
        return ops.window_signal(self._cuda(_x), self.win_size_samples, self.win_shift_samples, self.apply_padding_windowing)

    def frame_means(self, x: Tensor) -> Tensor:
        from ste_gan_b200 import ops
        return ops.frame_stats(self._cuda(x), self.win_size_samples, self.win_shift_samples, self.apply_padding_windowing, True, False)[0]

    def frame_power(self, x: Tensor) -> Tensor:
        from ste_gan_b200 import ops
        return ops.frame_stats(self._cuda(x), self.win_size_samples, self.win_shift_samples, self.apply_padding_windowing, False, True)[1]

    def double_average(self, x: Tensor) -> Tensor:
        """[B,T,C] -> avg(avg(x)) along time (time_domain_loss.py:51-55)."""
        xt = self._cuda(x).transpose(1, 2).contiguous()
        return self.average_filter(self.average_filter(xt)).transpose(1, 2).contiguous()

    def calculate_time_domain_features(self, raw_x: Tensor) -> Tensor:
        """[B,T,C] -> [B,F,C,4] (time_domain_loss.py:57-68)."""
        from ste_gan_b200 import ops
        return ops.td_features(self._cuda(raw_x), self.win_size_samples, self.win_shift_samples, self.apply_padding_windowing,
                               self.avg_filter_window_size)

    def time_domain_loss(self, x_real: Tensor, x_generated: Tensor):
        from ste_gan_b200.autograd import TdLossFn
        res = ((self.win_size_samples, self.win_shift_samples),)
        return TdLossFn.apply(x_real, x_generated, res, self.apply_padding_windowing, self.avg_filter_window_size)[0]

    def forward(self, x_real: Tensor, x_generated: Tensor) -> Tensor:
        return self.time_domain_loss(x_real, x_generated)


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
