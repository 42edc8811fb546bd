"""Convolution layers of the hot path: drop-in mirrors of ste_gan/layers/conv.py.

`WNConv1d`, `NormedConv1d`, `NormedConv2d`, `get_padding` and `GBlock` keep the
reference's names, constructor arguments, parameter names / shapes / registration
order (`bias`, `weight_g`, `weight_v` for weight-norm; `bias`, `weight_orig` + buffers
`weight_u`, `weight_v` for spectral-norm) and random initialisation (same RNG
consumption, so `torch.manual_seed(s)` gives bit-identical weights to the reference).

The modules hold parameters; the arithmetic runs in the CUDA library.  Calling a single
layer (`conv(x)` with the reference's [B,C,T] / [B,C,H,W] layout) goes through a
per-layer autograd function; the models call the fused whole-network passes instead
(ste_gan_b200/passes.py).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..precision import act_dtype


def get_padding(kernel_size, dilation=1):
    """layers/conv.py:24-25."""
    return int((kernel_size * dilation - dilation) / 2)


def _pair_first(v):
    return v[0] if isinstance(v, (tuple, list)) else v


class _NormedConvBase(nn.Module):
    """Parameter container for a conv1d, or a conv2d whose kernel / stride / padding /
    dilation act on the first spatial axis only ((k,1) kernels of the period discriminators)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 two_d=False):
        super().__init__()
        if two_d:
            for name, v in (("kernel_size", kernel_size), ("stride", stride), ("padding", padding), ("dilation", dilation)):
                if isinstance(v, (tuple, list)) and len(v) == 2:
                    second = v[1]
                    expect = 0 if name == "padding" else 1
                    if second != expect:
                        raise ValueError(f"{name}={v}: only (k,1)-shaped 2-D convolutions are on the hot path")
        if not bias:
            raise ValueError("bias=False is not used on the hot path")
        self.two_d = two_d
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel = _pair_first(kernel_size)
        self.stride, self.pad = _pair_first(stride), _pair_first(padding)
        self.dilation, self.groups = _pair_first(dilation), groups
        # reference initialisation: nn.Conv1d / nn.Conv2d reset_parameters (kaiming_uniform(a=sqrt 5) + uniform bias)
        proto = (nn.Conv2d(in_channels, out_channels, (self.kernel, 1), (self.stride, 1), (self.pad, 0),
                           (self.dilation, 1), groups) if two_d else
                 nn.Conv1d(in_channels, out_channels, self.kernel, self.stride, self.pad, self.dilation, groups))
        self._proto_weight = proto.weight.data
        self.bias = nn.Parameter(proto.bias.data)

    # -- shape helpers ------------------------------------------------------------
    def t_out(self, t_in: int) -> int:
        return (t_in + 2 * self.pad - self.dilation * (self.kernel - 1) - 1) // self.stride + 1

    def extra_repr(self) -> str:
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel}, stride={self.stride}, "
                f"padding={self.pad}, dilation={self.dilation}, groups={self.groups}, two_d={self.two_d}")

    # -- per-layer drop-in forward (reference layout) --------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ..passes import single_conv_forward
        return single_conv_forward(self, x)


class WeightNormConv(_NormedConvBase):
    """weight_norm(nn.Conv1d|nn.Conv2d) - layers/conv.py:16-17,92,99.  w = g * v / ||v||."""
    norm = "weight_norm"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        w = self._proto_weight
        del self._proto_weight
        dims = tuple(range(1, w.dim()))
        # torch.nn.utils.weight_norm: g = ||w|| over all dims but 0 (keepdim), v = w
        self.weight_g = nn.Parameter(torch.norm(w, 2, dim=dims, keepdim=True).data)
        self.weight_v = nn.Parameter(w)


class SpectralNormConv(_NormedConvBase):
    """spectral_norm(nn.Conv1d|nn.Conv2d) (legacy hook API) - layers/conv.py:94,101."""
    norm = "spectral_norm"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        w = self._proto_weight
        del self._proto_weight
        h, wd = w.shape[0], w[0].numel()
        # torch.nn.utils.spectral_norm.SpectralNorm.apply: u, v ~ normalize(N(0,1)), drawn after the conv init
        u = F.normalize(w.new_empty(h).normal_(0, 1), dim=0, eps=1e-12)
        v = F.normalize(w.new_empty(wd).normal_(0, 1), dim=0, eps=1e-12)
        self.weight_orig = nn.Parameter(w)
        self.register_buffer("weight_u", u)
        self.register_buffer("weight_v", v)


def WNConv1d(*args, **kwargs):
    """layers/conv.py:16-17."""
    return WeightNormConv(*args, **kwargs)


def NormedConv1d(*args, **kwargs):
    """layers/conv.py:89-94."""
    norm = kwargs.pop("norm", "weight_norm")
    if norm == "weight_norm":
        return WeightNormConv(*args, **kwargs)
    elif norm == "spectral_norm":
        return SpectralNormConv(*args, **kwargs)


def NormedConv2d(*args, **kwargs):
    """layers/conv.py:96-101."""
    norm = kwargs.pop("norm", "weight_norm")
    if norm == "weight_norm":
        return WeightNormConv(*args, two_d=True, **kwargs)
    elif norm == "spectral_norm":
        return SpectralNormConv(*args, two_d=True, **kwargs)


class GBlock(nn.Module):
    """layers/conv.py:29-84.  Same Sequential layout (ReLU / Upsample placeholders keep the
    indices of the parametrised layers identical: conv1.1/conv1.3/res1.0 or, with upsampling,
    conv1.2/conv1.4/res1.1; conv2.1/conv2.3)."""

    def __init__(self, input_dim, output_dim, upsample=1, kernel_size=3):
        super().__init__()
        self.input_dim, self.output_dim, self.upsample, self.kernel_size = input_dim, output_dim, upsample, kernel_size
        conv1 = [nn.ReLU()]
        if upsample > 1:
            conv1 += [nn.Upsample(scale_factor=upsample)]
        conv1 += [
            WNConv1d(input_dim, output_dim, kernel_size=kernel_size, padding=get_padding(kernel_size)),
            nn.ReLU(),
            WNConv1d(output_dim, output_dim, kernel_size=kernel_size, dilation=3, padding=get_padding(kernel_size, 3))]
        res1 = [nn.Upsample(scale_factor=upsample)] if upsample > 1 else []
        res1 += [WNConv1d(input_dim, output_dim, kernel_size=1)]
        conv2 = [
            nn.ReLU(),
            WNConv1d(output_dim, output_dim, kernel_size=kernel_size, dilation=9, padding=get_padding(kernel_size, 9)),
            nn.ReLU(),
            WNConv1d(output_dim, output_dim, kernel_size=kernel_size, dilation=27, padding=get_padding(kernel_size, 27))]
        self.conv1 = nn.Sequential(*conv1)
        self.res1 = nn.Sequential(*res1)
        self.conv2 = nn.Sequential(*conv2)

    # parametrised layers in execution order
    def convs(self):
        o = 1 if self.upsample > 1 else 0
        return dict(c1=self.conv1[1 + o], c2=self.conv1[3 + o], res=self.res1[o], c3=self.conv2[1], c4=self.conv2[3])

    def forward(self, x):
        """x [B, C_in, T] -> [B, C_out, upsample*T]; conv1(x) + res1(x), then x + conv2(x) (conv.py:82-84)."""
        from ..passes import gblock_forward_torch_layout
        return gblock_forward_torch_layout(self, x)
