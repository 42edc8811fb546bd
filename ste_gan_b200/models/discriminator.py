"""Discriminators - drop-in mirror of ste_gan/models/discriminator.py.

Same classes, constructor arguments, `.layers` / `.output` / `.period` / `.name` attributes,
`discriminator_names`, parameter names and random initialisation as the reference
(discriminator.py:19-203).  `DiscriminatorSmall` / `Discriminator` forward through the fused
CUDA pass; the sub-discriminators are callable on their own as well.
"""
import logging
from typing import List

import torch
import torch.nn as nn

from ste_gan_b200.layers.conv import NormedConv1d, NormedConv2d


class _SubDisc(nn.Module):
    kind = "S"

    def forward(self, x):
        """x: [B, C, T] (reference layout) -> list of feature maps, logits last."""
        from ste_gan_b200.autograd import DiscriminatorFn
        shim = _SingleShim(self)
        outs = DiscriminatorFn.apply(shim, x.transpose(1, 2), *self.parameters())
        return list(outs)


class _SingleShim:
    """Presents one sub-discriminator with the container interface the fused pass walks."""

    def __init__(self, sub):
        self.multi_pooled_disc = [sub] if sub.kind == "P" else []
        self.multi_scale_disc = [sub] if sub.kind == "S" else []
        self.training = sub.training
        self._sub = sub

    def parameters(self):
        return self._sub.parameters()


class DiscriminatorP(_SubDisc):
    """discriminator.py:19-43."""
    kind = "P"

    def __init__(self, num_emg_channels: int, period, norm="weight_norm", name="DiscriminatorP"):
        super().__init__()
        self.name = name
        self.num_emg_channels = num_emg_channels
        self.layers = nn.ModuleList([
            NormedConv2d(num_emg_channels, 32, (5, 1), (3, 1), padding=(2, 0), norm=norm),
            NormedConv2d(32, 128, (5, 1), (3, 1), padding=(2, 0), norm=norm),
            NormedConv2d(128, 512, (5, 1), (3, 1), padding=(2, 0), norm=norm),
            NormedConv2d(512, 1024, (5, 1), (3, 1), padding=(2, 0), norm=norm),
            NormedConv2d(1024, 1024, (5, 1), 1, padding=(2, 0), norm=norm)])
        self.output = NormedConv2d(1024, 1, kernel_size=(3, 1), padding=(1, 0))
        self.period = period


class DiscriminatorSmallerS(_SubDisc):
    """discriminator.py:47-67."""

    def __init__(self, num_emg_channels, norm="weight_norm", name="DiscriminatorS"):
        super().__init__()
        self.name = name
        self.num_emg_channels = num_emg_channels
        self.layers = nn.ModuleList([
            NormedConv1d(num_emg_channels, 128, 15, 1, padding=7, norm=norm),
            NormedConv1d(128, 256, 37, 2, groups=4, padding=18, norm=norm),
            NormedConv1d(256, 512, 37, 2, groups=16, padding=18, norm=norm),
            NormedConv1d(512, 1024, 5, 1, padding=2, norm=norm)])
        self.output = NormedConv1d(1024, 1, 3, 1, padding=1)


class DiscriminatorSmallerP(_SubDisc):
    """discriminator.py:70-93."""
    kind = "P"

    def __init__(self, num_emg_channels: int, period, norm="weight_norm", name="DiscriminatorP"):
        super().__init__()
        self.name = name
        self.num_emg_channels = num_emg_channels
        self.layers = nn.ModuleList([
            NormedConv2d(num_emg_channels, 32, (3, 1), (1, 1), padding=(2, 0), norm=norm),
            NormedConv2d(32, 256, (3, 1), (3, 1), padding=(2, 0), norm=norm),
            NormedConv2d(256, 512, (3, 1), (3, 1), padding=(2, 0), norm=norm),
        ])
        self.output = NormedConv2d(512, 1, kernel_size=(3, 1), padding=(1, 0))
        self.period = period


class DiscriminatorS(_SubDisc):
    """discriminator.py:96-119."""

    def __init__(self, num_emg_channels, norm="weight_norm", name="DiscriminatorS"):
        super().__init__()
        self.name = name
        self.num_emg_channels = num_emg_channels
        self.layers = nn.ModuleList([
            NormedConv1d(num_emg_channels, 128, 15, 1, padding=7, norm=norm),
            NormedConv1d(128, 128, 41, 2, groups=4, padding=20, norm=norm),
            NormedConv1d(128, 256, 41, 2, groups=16, padding=20, norm=norm),
            NormedConv1d(256, 512, 41, 4, groups=16, padding=20, norm=norm),
            NormedConv1d(512, 1024, 41, 4, groups=16, padding=20, norm=norm),
            NormedConv1d(1024, 1024, 41, 1, groups=16, padding=20, norm=norm),
            NormedConv1d(1024, 1024, 5, 1, padding=2, norm=norm)])
        self.output = NormedConv1d(1024, 1, 3, 1, padding=1)


class _MultiDisc(nn.Module):
    _P, _S = None, None

    def __init__(self, num_emg_channels: int, num_multi_pool=5, num_multi_scale=3):
        super().__init__()
        self.num_emg_channels = num_emg_channels
        prime_ratios = [2, 3, 5, 7, 11]
        self.multi_pooled_disc = nn.ModuleList([
            self._P(num_emg_channels, prime_ratios[i], name=f"DiscriminatorP-{prime_ratios[i]}")
            for i in range(num_multi_pool)])
        self.multi_scale_disc = nn.ModuleList([
            self._S(num_emg_channels=num_emg_channels, norm="spectral_norm" if i == 0 else "weight_norm",
                    name=f"DiscriminatorS-{i}")
            for i in range(num_multi_scale)])
        self.downsample = nn.AvgPool1d(kernel_size=4, stride=2, padding=1)   # arithmetic lives in stg_avgpool4
        self.discriminator_names = [d.name for d in self.multi_pooled_disc] + [d.name for d in self.multi_scale_disc]

    def forward(self, x) -> List[List[torch.Tensor]]:
        """x: [B, T, C] -> 8 lists of feature maps in the reference layouts, logits last."""
        from ste_gan_b200.autograd import DiscriminatorFn
        outs = DiscriminatorFn.apply(self, x, *self.parameters())
        results, i = [], 0
        for d in list(self.multi_pooled_disc) + list(self.multi_scale_disc):
            n = len(d.layers) + 1
            results.append(list(outs[i:i + n]))
            i += n
        return results


class DiscriminatorSmall(_MultiDisc):
    """discriminator.py:122-155."""
    _P, _S = DiscriminatorSmallerP, DiscriminatorSmallerS


class Discriminator(_MultiDisc):
    """discriminator.py:158-191."""
    _P, _S = DiscriminatorP, DiscriminatorS


def init_emg_discriminators(cfg) -> Discriminator:
    """discriminator.py:194-203."""
    num_emg_channels = cfg.data.num_emg_channels
    discriminator_small = cfg.model.discriminator_small
    if discriminator_small:
        logging.info(f"Initializing small discriminators with {num_emg_channels} channels")
        return DiscriminatorSmall(num_emg_channels)
    logging.info(f"Initializing FULL discriminators with {num_emg_channels} channels")
    return Discriminator(num_emg_channels)
