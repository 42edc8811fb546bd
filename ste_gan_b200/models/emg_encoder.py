"""EMG encoder - drop-in mirror of ste_gan/models/emg_encoder.py (+ the layers it is made of: ResBlock,
ste_gan/layers/conv.py:106-132, and TransformerEncoderLayer / MultiHeadAttention / LearnedRelativePositionalEmbedding,
ste_gan/layers/transformer.py).

Same classes, constructor arguments, parameter / buffer names, registration order and random initialisation as the
reference, so a reference checkpoint (`load_emg_encoder`, emg_encoder.py:113-124) loads key for key.  The torch modules used
inside (nn.Conv1d, nn.BatchNorm1d, nn.Linear, nn.LayerNorm) are PARAMETER CONTAINERS only - their forward is never called:
the arithmetic is the fused CUDA pass of ste_gan_b200/passes_encoder.py (ResBlocks with eval-mode BatchNorm folded into
weights and bias + every projection on the tcgen05 convolution engine, attention / LayerNorm / losses in csrc/encoder.cu).

Only what the GAN train step uses is on this path: the FROZEN encoder in eval mode
(losses/emg_encoder_loss.py:61), forward and the gradient w.r.t. the EMG input.  Training the encoder itself
(ste_gan/emg_encoder/train.py: random input shift, dropout, BatchNorm batch statistics) is a different program and out of
scope; `forward` raises in training mode.
"""
from __future__ import annotations

import copy
from pathlib import Path
from typing import Tuple

import torch
import torch.nn as nn
from torch import Tensor

import ste_gan_b200 as ste_gan

PHONEME_INVENTORY = ['aa', 'ae', 'ah', 'ao', 'aw', 'ax', 'axr', 'ay', 'b', 'ch', 'd', 'dh', 'dx', 'eh', 'el', 'em', 'en', 'er', 'ey',
                     'f', 'g', 'hh', 'hv', 'ih', 'iy', 'jh', 'k', 'l', 'm', 'n', 'nx', 'ng', 'ow', 'oy', 'p', 'r', 's', 'sh', 't', 'th',
                     'uh', 'uw', 'v', 'w', 'y', 'z', 'zh', 'sil']           # constants.py:148
SILENCE_PHONEME_INDEX = PHONEME_INVENTORY.index("sil")                  # constants.py:150


class ResBlock(nn.Module):
    """layers/conv.py:106-132: conv(k3, stride) - BN - ReLU - conv(k3) - BN, + (1x1 strided conv - BN) residual, ReLU."""

    def __init__(self, num_ins, num_outs, stride=1):
        super().__init__()
        self.conv1 = nn.Conv1d(num_ins, num_outs, 3, padding=1, stride=stride)
        self.bn1 = nn.BatchNorm1d(num_outs)
        self.conv2 = nn.Conv1d(num_outs, num_outs, 3, padding=1)
        self.bn2 = nn.BatchNorm1d(num_outs)
        if stride != 1 or num_ins != num_outs:
            self.residual_path = nn.Conv1d(num_ins, num_outs, 1, stride=stride)
            self.res_norm = nn.BatchNorm1d(num_outs)
        else:
            self.residual_path = None
        self.stride = stride

    def forward(self, x):
        raise RuntimeError("ResBlock is a parameter container here: use EMGEncoderTransformer.forward (fused CUDA pass)")


class LearnedRelativePositionalEmbedding(nn.Module):
    """layers/transformer.py:115-171 (parameters only; the logits are computed inside stg_relattn_fwd)."""

    def __init__(self, max_relative_pos: int, num_heads: int, embedding_dim: int, unmasked: bool = False,
                 heads_share_embeddings: bool = False, add_to_values: bool = False):
        super().__init__()
        if not unmasked or heads_share_embeddings or add_to_values:
            raise ValueError("only the encoder's configuration (unmasked, per-head, keys only) is on this path")
        self.max_relative_pos, self.num_heads, self.embedding_dim = max_relative_pos, num_heads, embedding_dim
        self.embeddings = nn.Parameter(torch.zeros(num_heads, 2 * max_relative_pos - 1, embedding_dim, 1))
        nn.init.normal_(self.embeddings, mean=0.0, std=embedding_dim ** (-0.5))


class MultiHeadAttention(nn.Module):
    """layers/transformer.py:63-113."""

    def __init__(self, d_model=256, n_head=4, dropout=0.1, relative_positional=True, relative_positional_distance=100):
        super().__init__()
        self.d_model, self.n_head = d_model, n_head
        d_qkv = d_model // n_head
        assert d_qkv * n_head == d_model, 'd_model must be divisible by n_head'
        self.d_qkv = d_qkv
        self.w_q = nn.Parameter(torch.Tensor(n_head, d_model, d_qkv))
        self.w_k = nn.Parameter(torch.Tensor(n_head, d_model, d_qkv))
        self.w_v = nn.Parameter(torch.Tensor(n_head, d_model, d_qkv))
        self.w_o = nn.Parameter(torch.Tensor(n_head, d_qkv, d_model))
        nn.init.xavier_normal_(self.w_q)
        nn.init.xavier_normal_(self.w_k)
        nn.init.xavier_normal_(self.w_v)
        nn.init.xavier_normal_(self.w_o)
        self.dropout = nn.Dropout(dropout)
        if not relative_positional:
            raise ValueError("the encoder uses relative positional logits (emg_encoder.py:62-65)")
        self.relative_positional = LearnedRelativePositionalEmbedding(relative_positional_distance, n_head, d_qkv, True)


class TransformerEncoderLayer(nn.Module):
    """layers/transformer.py:8-60 (post-norm: x = LN(x + attn(x)); x = LN(x + W2 relu(W1 x)))."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, relative_positional=True,
                 relative_positional_distance=100):
        super().__init__()
        self.self_attn = MultiHeadAttention(d_model, nhead, dropout=dropout, relative_positional=relative_positional,
                                            relative_positional_distance=relative_positional_distance)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)


class _EncoderStack(nn.Module):
    """nn.TransformerEncoder(encoder_layer, num_layers) as the reference builds it (emg_encoder.py:66): `num_layers` deep
    copies of one initialised layer, registered as `layers.{i}`; no final norm, no masks."""

    def __init__(self, encoder_layer: nn.Module, num_layers: int):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers


class EMGEncoder(nn.Module):
    """Base class for EMG encoders (emg_encoder.py:25-33)."""

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        raise NotImplementedError("Must be implemented by subclasses.")


class EMGEncoderTransformer(EMGEncoder):
    """The Conv-Transformer EMG encoder (emg_encoder.py:36-88)."""

    def __init__(self, num_ins, num_outs, num_aux_outs, model_size: int = 768, num_extra_res_blocks: int = 3,
                 dropout: float = 0.2, num_transformer_layers: int = 6):
        super().__init__()
        res_blocks = [ResBlock(num_ins, model_size, 2)]
        for _ in range(num_extra_res_blocks):
            res_blocks.append(ResBlock(model_size, model_size, 2))
        self.conv_blocks = nn.Sequential(*res_blocks)
        self.w_raw_in = nn.Linear(model_size, model_size)
        encoder_layer = TransformerEncoderLayer(d_model=model_size, nhead=8, relative_positional=True,
                                                relative_positional_distance=100, dim_feedforward=3072, dropout=dropout)
        self.transformer = _EncoderStack(encoder_layer, num_transformer_layers)
        self.w_out = nn.Linear(model_size, num_outs)
        self.w_aux = nn.Linear(model_size, num_aux_outs)
        self.model_size = model_size
        self._plan = None            # packed frozen operands (passes_encoder.EncoderPlan), built on first use

    def invalidate_plan(self) -> None:
        """Call after the weights changed (load_state_dict does it): the packed operands are rebuilt on the next forward."""
        self._plan = None

    def load_state_dict(self, *a, **kw):
        self._plan = None
        return super().load_state_dict(*a, **kw)

    def plan(self, dtype: torch.dtype):
        from ste_gan_b200.passes_encoder import EncoderPlan
        if self._plan is None or self._plan.dtype != dtype or self._plan.device != self.w_out.weight.device:
            self._plan = EncoderPlan(self, dtype)
        return self._plan

    def forward(self, x_raw: Tensor) -> Tuple[Tensor, Tensor]:
        """[B, T, C] EMG -> (speech-unit prediction [B, T/16, num_outs], phoneme logits [B, T/16, num_aux_outs]).
        Eval mode only (the frozen encoder of the GAN step); differentiable w.r.t. x_raw."""
        if self.training:
            raise RuntimeError("EMGEncoderTransformer: only the frozen eval-mode encoder is on this path - call .eval() "
                               "(encoder training, ste_gan/emg_encoder/train.py, is out of scope)")
        from ste_gan_b200.autograd import EncoderFn
        return EncoderFn.apply(self, x_raw)


def init_emg_encoder(cfg, device: torch.device = None) -> EMGEncoder:
    """emg_encoder.py:91-111."""
    emg_encoder_config = cfg.emg_encoder
    num_ins: int = cfg.data.num_emg_channels
    num_outs: int = ste_gan.SPEECH_UNITS_FEAT_SIZE
    num_aux_outs: int = len(PHONEME_INVENTORY)
    emg_encoder_type = emg_encoder_config["type"]
    emg_encoder_params = emg_encoder_config["params"]
    emg_encoder_args = dict(num_ins=num_ins, num_outs=num_outs, num_aux_outs=num_aux_outs)
    if emg_encoder_type == "EMGEncoderTransformer":
        emg_encoder = EMGEncoderTransformer(**emg_encoder_args, **emg_encoder_params)
    else:
        raise ValueError(f"Unknown EMG encoder type: {emg_encoder_type}")
    if device:
        emg_encoder = emg_encoder.to(device)
    return emg_encoder


def load_emg_encoder(cfg, device: torch.device, emg_encoder_checkpoint_path: Path) -> EMGEncoder:
    """emg_encoder.py:113-124."""
    emg_encoder = init_emg_encoder(cfg, device)
    state_dict = torch.load(emg_encoder_checkpoint_path, map_location=device)
    emg_encoder.load_state_dict(state_dict)
    emg_encoder.eval()
    return emg_encoder
