"""Arithmetic mode of the drop-in modules.

  "fp32" - validation mode: fp32 storage, CUDA-core FFMA kernels (relative L2 <= 1e-4 vs the reference).
  "bf16" - production mode: bf16 activations / packed weights, tcgen05 tensor cores, fp32 accumulation,
           fp32 master weights, reductions and losses (relative L2 <= 2e-2).
  None   - follow `torch.autocast`: bf16 inside an enabled CUDA autocast region (the reference's
           `train.mixed_precision: true`, ste_gan/train.py:181,204 - fp16 there, bf16 here), else fp32.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

_mode: Optional[str] = None


def set_precision(mode: Optional[str]) -> None:
    global _mode
    if mode not in (None, "fp32", "bf16"):
        raise ValueError(f"unknown precision {mode!r}")
    _mode = mode


def get_precision() -> str:
    if _mode is not None:
        return _mode
    return "bf16" if torch.is_autocast_enabled("cuda") else "fp32"


def act_dtype() -> torch.dtype:
    return torch.bfloat16 if get_precision() == "bf16" else torch.float32


@contextlib.contextmanager
def precision(mode: Optional[str]):
    global _mode
    old = _mode
    set_precision(mode)
    try:
        yield
    finally:
        _mode = old
