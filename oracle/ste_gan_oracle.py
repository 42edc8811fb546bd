"""CPU oracle for the STE-GAN hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *functional restatement* (plain torch fp32/fp64 CPU ops, no
nn.Module from the reference, no CUDA) of the reference algorithm on the hot
path: generator, discriminator stacks, multi time-domain loss, LSGAN and
feature-matching losses and the train-step sequencing.  Every function cites
the reference file:line (relative to /root/reference) it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this module, and only as the checker or the
reported CPU baseline - never on the product path.

Pinning: the reference repository ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
modules themselves, imported from /root/reference in the authoring container by
`oracle/make_golden.py`, which writes `tests/golden/*.pt`.
`tests/test_oracle_golden.py` checks the oracle against those fixtures (and,
where /root/reference is present, against the live reference modules).

Everything is keyed by the reference's own `state_dict` names
(`gblocks.3.conv1.2.weight_v`, `multi_scale_disc.0.layers.1.weight_orig`, ...)
because those names are the drop-in contract.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

# ste_gan/constants.py:26-38
SPEECH_UNITS_FEAT_SIZE = 256
NUM_MFCCS = 25
EMBEDDING_DIM_SIZE = 64
NUM_EMG_CHANNELS = 8
NUM_EMG_SESSIONS = 17
PRIME_RATIOS = [2, 3, 5, 7, 11]  # models/discriminator.py:128,164


# --------------------------------------------------------------------------
# weight re-parametrisations
# --------------------------------------------------------------------------
def weight_norm_weight(g: Tensor, v: Tensor) -> Tensor:
    """`torch.nn.utils.weight_norm(dim=0)` as applied by layers/conv.py:16-17,92,99:
    w = g * v / ||v||, the norm taken over every dim but 0."""
    dims = tuple(range(1, v.dim()))
    return v * (g / v.norm(2, dim=dims, keepdim=True))


def spectral_norm_weight(w_orig: Tensor, u: Tensor, v: Tensor, training: bool,
                         eps: float = 1e-12) -> Tensor:
    """Legacy `torch.nn.utils.spectral_norm` (layers/conv.py:94,101): one power
    iteration per *training-mode forward call*, performed IN PLACE on the
    buffers `u` and `v`; sigma = u^T W v with u, v treated as constants."""
    wm = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=eps))
    uc, vc = u.detach().clone(), v.detach().clone()
    sigma = torch.dot(uc, torch.mv(wm, vc))
    return w_orig / sigma


def _normed_weight(sd: StateDict, prefix: str, training: bool) -> Tensor:
    """NormedConv1d/2d (layers/conv.py:89-101): weight-norm or spectral-norm keys."""
    if prefix + "weight_g" in sd:
        return weight_norm_weight(sd[prefix + "weight_g"], sd[prefix + "weight_v"])
    return spectral_norm_weight(sd[prefix + "weight_orig"], sd[prefix + "weight_u"],
                                sd[prefix + "weight_v"], training)


# --------------------------------------------------------------------------
# activations with an injected sign pattern (parity tests only)
# --------------------------------------------------------------------------
class _ActWithMask(torch.autograd.Function):
    """relu / leaky_relu whose BACKWARD takes the slope pattern from `mask` (True: slope 1) instead of from
    the sign of its own input.  ReLU / LeakyReLU have a discontinuous derivative: a pre-activation that is
    zero to within rounding takes the other slope in a lower-precision implementation and changes every
    gradient upstream of it by a finite amount.  Feeding the implementation-under-test's OWN sign pattern
    into the oracle's backward makes both differentiate the same piecewise-linear branch, so the remaining
    gradient difference is arithmetic only and every tensor can be held to the stated tolerance."""

    @staticmethod
    def forward(ctx, x, mask, slope):
        ctx.save_for_backward(mask)
        ctx.slope = slope
        return torch.where(x > 0, x, x * slope)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * torch.where(mask, torch.ones((), dtype=g.dtype), torch.full((), ctx.slope, dtype=g.dtype)), None, None


def _act(x: Tensor, slope: float, mask: Optional[Tensor]) -> Tensor:
    """relu (slope 0) / leaky_relu(slope); with `mask` (bool, same shape) the backward uses its pattern."""
    if mask is None:
        return F.relu(x) if slope == 0.0 else F.leaky_relu(x, slope)
    assert mask.shape == x.shape, (mask.shape, x.shape)
    return _ActWithMask.apply(x, mask, slope)


# --------------------------------------------------------------------------
# generator (models/generator.py:78-162, layers/conv.py:29-84)
# --------------------------------------------------------------------------
def _wnconv1d(sd: StateDict, prefix: str, x: Tensor, dilation: int = 1, padding: int = 0) -> Tensor:
    w = weight_norm_weight(sd[prefix + "weight_g"], sd[prefix + "weight_v"])
    return F.conv1d(x, w, sd[prefix + "bias"], dilation=dilation, padding=padding)


def gblock_forward(sd: StateDict, prefix: str, x: Tensor, upsample: int,
                   masks: Optional[Sequence[Tensor]] = None) -> Tensor:
    """GBlock.forward (layers/conv.py:82-84).  Sequential indices shift by one
    when an nn.Upsample is present (layers/conv.py:38-56).
    masks (parity tests): sign patterns of the block's four ReLU inputs, see _ActWithMask."""
    o = 1 if upsample > 1 else 0
    up = (lambda t: F.interpolate(t, scale_factor=float(upsample), mode="nearest")) if upsample > 1 else (lambda t: t)
    m = masks if masks is not None else [None] * 4
    # conv1 = ReLU -> [Up] -> WNConv(k3,d1,p1) -> ReLU -> WNConv(k3,d3,p3)   (conv.py:38-53)
    h = _wnconv1d(sd, f"{prefix}conv1.{1 + o}.", up(_act(x, 0.0, m[0])), dilation=1, padding=1)
    h = _wnconv1d(sd, f"{prefix}conv1.{3 + o}.", _act(h, 0.0, m[1]), dilation=3, padding=3)
    # res1 = [Up] -> WNConv(k1) on the un-ReLU'd input                         (conv.py:55-56)
    r = _wnconv1d(sd, f"{prefix}res1.{o}.", up(x))
    h = h + r                                                                 # conv.py:83
    # conv2 = ReLU -> WNConv(k3,d9,p9) -> ReLU -> WNConv(k3,d27,p27)           (conv.py:61-75)
    c = _wnconv1d(sd, f"{prefix}conv2.1.", _act(h, 0.0, m[2]), dilation=9, padding=9)
    c = _wnconv1d(sd, f"{prefix}conv2.3.", _act(c, 0.0, m[3]), dilation=27, padding=27)
    return h + c                                                              # conv.py:84


def generator_upsamples(speech_feature_type: str = "SPEECH_UNITS") -> List[int]:
    """Per-GBlock upsample factors (models/generator.py:116-131)."""
    last = 2 if speech_feature_type == "SPEECH_UNITS" else 1
    return [1, 1, 2, 2, 2, last, 1, 1]


def generator_forward(sd: StateDict, speech_units: Tensor, session_ids: Tensor,
                      speaking_mode_ids: Optional[Tensor] = None,
                      speech_feature_type: str = "SPEECH_UNITS",
                      masks: Optional[Sequence] = None) -> Tensor:
    """EMGGeneratorGanTTS.forward (models/generator.py:140-162):
    [B,T,D] units (+ session / speaking-mode embeddings) -> [B,16T,8] in (-1,1).
    masks (parity tests): [4 masks per GBlock] * 8 + [mask of last_conv's ReLU], [B,C,T] bool."""
    x = speech_units
    T = x.shape[1]
    if "session_embeddings.weight" in sd:                                     # generator.py:143-146
        e = sd["session_embeddings.weight"][session_ids.long()]
        x = torch.cat((x, e[:, None, :].expand(-1, T, -1)), dim=-1)
    if "speaking_mode_embeddings.weight" in sd:                               # generator.py:148-151
        e = sd["speaking_mode_embeddings.weight"][speaking_mode_ids.long()]
        x = torch.cat((x, e[:, None, :].expand(-1, T, -1)), dim=-1)
    x = x.transpose(1, 2)                                                     # generator.py:154
    x = _wnconv1d(sd, "gblocks.0.", x)                                        # generator.py:119
    for i, up in enumerate(generator_upsamples(speech_feature_type)):         # generator.py:121-130
        x = gblock_forward(sd, f"gblocks.{i + 1}.", x, up, masks[i] if masks is not None else None)
    x = _wnconv1d(sd, "last_conv.1.", _act(x, 0.0, masks[-1] if masks is not None else None), padding=1)   # generator.py:133-137
    return torch.tanh(x.transpose(1, 2))                                      # generator.py:157-160


# --------------------------------------------------------------------------
# discriminators (models/discriminator.py)
# --------------------------------------------------------------------------
# (C_in, C_out, k, stride, pad, groups) per layer.
SMALL_P_LAYERS = lambda c: [(c, 32, 3, 1, 2, 1), (32, 256, 3, 3, 2, 1), (256, 512, 3, 3, 2, 1)]   # discriminator.py:76-80
FULL_P_LAYERS = lambda c: [(c, 32, 5, 3, 2, 1), (32, 128, 5, 3, 2, 1), (128, 512, 5, 3, 2, 1),     # discriminator.py:25-30
                           (512, 1024, 5, 3, 2, 1), (1024, 1024, 5, 1, 2, 1)]
SMALL_S_LAYERS = lambda c: [(c, 128, 15, 1, 7, 1), (128, 256, 37, 2, 18, 4),                      # discriminator.py:54-58
                            (256, 512, 37, 2, 18, 16), (512, 1024, 5, 1, 2, 1)]
FULL_S_LAYERS = lambda c: [(c, 128, 15, 1, 7, 1), (128, 128, 41, 2, 20, 4), (128, 256, 41, 2, 20, 16),  # discriminator.py:103-110
                           (256, 512, 41, 4, 20, 16), (512, 1024, 41, 4, 20, 16),
                           (1024, 1024, 41, 1, 20, 16), (1024, 1024, 5, 1, 2, 1)]


def disc_p_forward(sd: StateDict, prefix: str, x: Tensor, period: int, small: bool, training: bool,
                   masks: Optional[Sequence[Tensor]] = None) -> List[Tensor]:
    """DiscriminatorSmallerP.forward / DiscriminatorP.forward (discriminator.py:84-93 / :34-43).
    x is [B,C,T].  Reflect pad on the right by period - T % period (always >= 1)."""
    layers = (SMALL_P_LAYERS if small else FULL_P_LAYERS)(x.shape[1])
    x = F.pad(x, (0, period - x.shape[-1] % period), "reflect")
    x = x.view(x.shape[0], x.shape[1], x.shape[2] // period, period)
    fmaps = []
    for j, (_, _, k, s, p, _) in enumerate(layers):
        w = _normed_weight(sd, f"{prefix}layers.{j}.", training)
        x = _act(F.conv2d(x, w, sd[f"{prefix}layers.{j}.bias"], stride=(s, 1), padding=(p, 0)), 0.1,
                 masks[j] if masks is not None else None)
        fmaps.append(x)
    w = _normed_weight(sd, f"{prefix}output.", training)
    fmaps.append(F.conv2d(x, w, sd[f"{prefix}output.bias"], padding=(1, 0)))
    return fmaps


def disc_s_forward(sd: StateDict, prefix: str, x: Tensor, small: bool, training: bool,
                   masks: Optional[Sequence[Tensor]] = None) -> List[Tensor]:
    """DiscriminatorSmallerS.forward / DiscriminatorS.forward (discriminator.py:61-67 / :113-119)."""
    layers = (SMALL_S_LAYERS if small else FULL_S_LAYERS)(x.shape[1])
    fmaps = []
    for j, (_, _, k, s, p, g) in enumerate(layers):
        w = _normed_weight(sd, f"{prefix}layers.{j}.", training)
        x = _act(F.conv1d(x, w, sd[f"{prefix}layers.{j}.bias"], stride=s, padding=p, groups=g), 0.1,
                 masks[j] if masks is not None else None)
        fmaps.append(x)
    w = _normed_weight(sd, f"{prefix}output.", training)
    fmaps.append(F.conv1d(x, w, sd[f"{prefix}output.bias"], padding=1))
    return fmaps


def discriminator_forward(sd: StateDict, x: Tensor, small: bool = True, training: bool = True,
                          num_multi_pool: int = 5, num_multi_scale: int = 3,
                          masks: Optional[Sequence[Sequence[Tensor]]] = None) -> List[List[Tensor]]:
    """DiscriminatorSmall.forward / Discriminator.forward (discriminator.py:144-155 / :180-191).
    x is [B,T,C]; returns 8 lists of feature maps with the logits last.
    masks (parity tests): per sub-discriminator, the sign pattern of every LeakyReLU output (reference layout)."""
    x = x.transpose(1, 2)
    results = []
    mk = lambda i: masks[i] if masks is not None else None
    for i in range(num_multi_pool):
        results.append(disc_p_forward(sd, f"multi_pooled_disc.{i}.", x, PRIME_RATIOS[i], small, training, mk(i)))
    for i in range(num_multi_scale):
        results.append(disc_s_forward(sd, f"multi_scale_disc.{i}.", x, small, training, mk(num_multi_pool + i)))
        x = F.avg_pool1d(x, kernel_size=4, stride=2, padding=1)              # discriminator.py:140,153
    return results


# --------------------------------------------------------------------------
# multi time-domain feature loss (losses/time_domain_loss.py, layers/average_filter.py)
# --------------------------------------------------------------------------
TD_RESOLUTIONS = [(20, 8), (51, 13), (80, 16)]        # time_domain_loss.py:88-93 (win, shift)


def average_filter(x: Tensor, window: int = 9) -> Tensor:
    """AverageFilter.forward (average_filter.py:22-28): reflect pad w//2, AvgPool1d(w, stride 1). x is [B,C,T]."""
    p = window // 2
    return F.avg_pool1d(F.pad(x, (p, p), mode="reflect"), kernel_size=window, stride=1)


def window_signal(s: Tensor, win: int, shift: int, pad: bool = True) -> Tensor:
    """TimeDomainFeatureLoss.window_signal (time_domain_loss.py:35-41): [B,T,C] -> [B,F,C,win].
    F.pad(x, (0,0,p,p), 'reflect') on a 3-D tensor pads the time axis (dim 1)."""
    if pad:
        p = win // 2
        idx = torch.arange(-p, s.shape[1] + p).abs()
        idx = torch.where(idx >= s.shape[1], 2 * (s.shape[1] - 1) - idx, idx)
        s = s[:, idx, :]
    return s.unfold(1, win, shift)


def td_features(x: Tensor, win: int, shift: int, pad: bool = True, avg_window: int = 9) -> Tensor:
    """calculate_time_domain_features (time_domain_loss.py:57-68). x is [B,T,C] -> [B,F,C,4]."""
    low = average_filter(average_filter(x.transpose(1, 2), avg_window), avg_window).transpose(1, 2)   # :51-55
    hi = (x - low).abs()                                                      # :59-60
    fl, fh = window_signal(low, win, shift, pad), window_signal(hi, win, shift, pad)
    return torch.stack([fl.mean(-1), (fl ** 2).sum(-1), (fh ** 2).sum(-1), fh.mean(-1)], dim=-1)  # :62-67


def td_loss(x_real: Tensor, x_gen: Tensor, win: int, shift: int, pad: bool = True, avg_window: int = 9) -> Tensor:
    """TimeDomainFeatureLoss.time_domain_loss (time_domain_loss.py:70-73)."""
    return F.l1_loss(td_features(x_gen, win, shift, pad, avg_window), td_features(x_real, win, shift, pad, avg_window).detach())


def multi_td_loss(x_real: Tensor, x_gen: Tensor) -> Tuple[Tensor, List[Tensor]]:
    """MultiTimeDomainFeatureLoss.time_domain_loss (time_domain_loss.py:96-103); note (real, generated) order."""
    parts = [td_loss(x_real, x_gen, w, s) for (w, s) in TD_RESOLUTIONS]
    return sum(parts), parts


# --------------------------------------------------------------------------
# adversarial + feature-matching losses (train.py:189-198, 205-211, 257-264)
# --------------------------------------------------------------------------
def lsgan_d_loss(d_fake: Sequence[Sequence[Tensor]], d_real: Sequence[Sequence[Tensor]]) -> Tensor:
    loss = 0
    for scale in d_fake:
        loss = loss + F.mse_loss(scale[-1], torch.zeros_like(scale[-1]))      # train.py:193-194
    for scale in d_real:
        loss = loss + F.mse_loss(scale[-1], torch.ones_like(scale[-1]))       # train.py:195-196
    return loss


def lsgan_g_loss(d_fake: Sequence[Sequence[Tensor]]) -> Tensor:
    loss = 0
    for scale in d_fake:
        loss = loss + F.mse_loss(scale[-1], torch.ones_like(scale[-1]))       # train.py:210-211
    return loss


def feature_matching_loss(d_fake, d_real) -> Tensor:
    loss = 0
    for i in range(len(d_fake)):
        for j in range(len(d_fake[i]) - 1):
            loss = loss + F.l1_loss(d_fake[i][j], d_real[i][j].detach())      # train.py:259-262
    return loss


# --------------------------------------------------------------------------
# one GAN train step (train.py:165-268), fp32, no AMP, no encoder losses
# --------------------------------------------------------------------------
W_TD, W_FM = 15.0, 7.0                                # configs/ste_gan_base_gantts.yaml:33,37


def _leaf(sd: StateDict, buffers: Sequence[str] = ("weight_u",)) -> StateDict:
    out = {}
    for k, v in sd.items():
        is_buf = k.endswith("weight_u") or (k.endswith("weight_v") and k[:-1] + "u" in sd)
        out[k] = v.detach().clone() if is_buf else v.detach().clone().requires_grad_(True)
    return out


def is_buffer_key(sd: StateDict, k: str) -> bool:
    """spectral-norm `weight_u` / `weight_v` are buffers, everything else a parameter."""
    return k.endswith("weight_u") or (k.endswith("weight_v") and (k[:-1] + "u") in sd)


def losses_and_grads(sd_g: StateDict, sd_d: StateDict, speech_units: Tensor, session_ids: Tensor,
                     x_real: Tensor, small: bool = True, speech_feature_type: str = "SPEECH_UNITS",
                     d_lr_step=None, masks: Optional[Dict[str, object]] = None,
                     speaking_mode_ids: Optional[Tensor] = None, encoder: Optional[Dict[str, object]] = None) -> Dict[str, object]:
    """One iteration of train.py:165-268 up to (and excluding) the optimizer
    arithmetic, returning every consumed quantity.  `d_lr_step(sd_d, grads)` -
    if given - is applied between the D and G phases (train.py:199) so that
    the G phase sees the updated discriminator, as in the reference.
    Spectral-norm buffers in `sd_d` advance 4 times (4 training forwards).
    masks (parity tests, see _ActWithMask): optional sign patterns under the keys "g" (generator_forward),
    "d_fake_det", "d_real", "d_fake" (discriminator_forward of that pass).
    encoder: optional dict(sd=<EMG-encoder state_dict>, phoneme_targets=[B, frames] int64, w_su=1.0, w_ph=1.0) - adds the
    two perceptual losses of train.py:219-230 (oracle/emg_encoder_oracle.py) to the generator loss."""
    masks = masks or {}
    g = {k: (v.detach().clone().requires_grad_(True)) for k, v in sd_g.items()}
    d = {k: (v if is_buffer_key(sd_d, k) else v.detach().clone().requires_grad_(True)) for k, v in sd_d.items()}
    out: Dict[str, object] = {}
    x_pred = generator_forward(g, speech_units, session_ids, speaking_mode_ids, speech_feature_type, masks.get("g"))   # train.py:182
    out["x_pred"] = x_pred.detach()
    d_fake_det = discriminator_forward(d, x_pred.detach(), small, masks=masks.get("d_fake_det"))   # :190
    d_real = discriminator_forward(d, x_real, small, masks=masks.get("d_real"))              # :191
    loss_d = lsgan_d_loss(d_fake_det, d_real)                                               # :192-196
    dparams = [k for k in d if not is_buffer_key(sd_d, k)]
    dgrads = torch.autograd.grad(loss_d, [d[k] for k in dparams])                           # :198
    out["loss_d"] = loss_d.detach()
    out["grad_d"] = dict(zip(dparams, dgrads))
    out["d_fake_det"] = [[f.detach() for f in fm] for fm in d_fake_det]
    out["d_real"] = [[f.detach() for f in fm] for fm in d_real]
    if d_lr_step is not None:                                                               # :199
        with torch.no_grad():
            d_lr_step(d, out["grad_d"])
    d_fake = discriminator_forward(d, x_pred, small, masks=masks.get("d_fake"))              # :206
    d_real2 = discriminator_forward(d, x_real, small)                                       # :207
    loss_adv = lsgan_g_loss(d_fake)                                                         # :209-211
    td, td_parts = multi_td_loss(x_real, x_pred)                                            # :215
    fm = feature_matching_loss(d_fake, d_real2)                                             # :257-262
    loss_g = loss_adv + W_TD * td + W_FM * fm                                               # :216,263
    if encoder is not None:                                                                 # :219-230
        from oracle import emg_encoder_oracle as E
        units, logits = E.emg_encoder_forward(encoder["sd"], x_pred, masks.get("encoder"))
        su_loss, ph_loss = E.encoder_losses(units, logits, speech_units, encoder["phoneme_targets"])
        loss_g = loss_g + encoder.get("w_su", 1.0) * su_loss + encoder.get("w_ph", 1.0) * ph_loss
        out.update(loss_speech_unit=su_loss.detach(), loss_phoneme=ph_loss.detach())
    gparams = list(g.keys())
    x_pred.retain_grad()
    ggrads = torch.autograd.grad(loss_g, [g[k] for k in gparams] + [x_pred])                # :266
    out.update(loss_g=loss_g.detach(), loss_adv=loss_adv.detach(), loss_td=td.detach(),
               loss_td_parts=[p.detach() for p in td_parts], loss_fm=fm.detach(),
               grad_g=dict(zip(gparams, ggrads[:-1])), grad_x_pred=ggrads[-1],
               d_fake=[[f.detach() for f in fm_] for fm_ in d_fake])
    return out


class OracleTrainer:
    """Restated train loop body with the two AdamW optimisers (constants.py:57,
    train.py:80-81,198-199,266-267) - used as the CPU baseline ("port") and for
    multi-step parity of the fused optimiser kernel."""

    def __init__(self, sd_g: StateDict, sd_d: StateDict, small: bool = True,
                 speech_feature_type: str = "SPEECH_UNITS"):
        self.small, self.sft = small, speech_feature_type
        self.g = {k: v.detach().clone().requires_grad_(True) for k, v in sd_g.items()}
        self.d = {k: (v.detach().clone() if is_buffer_key(sd_d, k) else v.detach().clone().requires_grad_(True))
                  for k, v in sd_d.items()}
        self.d_params = [k for k in self.d if not is_buffer_key(sd_d, k)]
        self.opt_g = torch.optim.AdamW(list(self.g.values()), lr=2e-4, betas=(0.8, 0.99))
        self.opt_d = torch.optim.AdamW([self.d[k] for k in self.d_params], lr=2e-4, betas=(0.8, 0.99))

    def step(self, speech_units: Tensor, session_ids: Tensor, x_real: Tensor) -> Dict[str, float]:
        self.opt_d.zero_grad(); self.opt_g.zero_grad()                                   # train.py:166-167
        x_pred = generator_forward(self.g, speech_units, session_ids, None, self.sft)    # :182
        d_fake_det = discriminator_forward(self.d, x_pred.detach(), self.small)          # :190
        d_real = discriminator_forward(self.d, x_real, self.small)                       # :191
        loss_d = lsgan_d_loss(d_fake_det, d_real)
        loss_d.backward(); self.opt_d.step()                                             # :198-199
        d_fake = discriminator_forward(self.d, x_pred, self.small)                       # :206
        d_real = discriminator_forward(self.d, x_real, self.small)                       # :207
        loss_adv = lsgan_g_loss(d_fake)
        td, _ = multi_td_loss(x_real, x_pred)
        fm = feature_matching_loss(d_fake, d_real)
        loss_g = loss_adv + W_TD * td + W_FM * fm
        loss_g.backward(); self.opt_g.step()                                             # :266-267
        return dict(loss_d=float(loss_d), loss_g=float(loss_g), loss_adv=float(loss_adv),
                    loss_td=float(td), loss_fm=float(fm))


# --------------------------------------------------------------------------
# synthetic inputs: shared by tests, smoke and bench
# --------------------------------------------------------------------------
def synthetic_batch(batch: int, frames: int, seed: int = 0, channels: int = NUM_EMG_CHANNELS,
                    unit_dim: int = SPEECH_UNITS_FEAT_SIZE, hop: int = 16,
                    num_sessions: int = NUM_EMG_SESSIONS):
    gen = torch.Generator().manual_seed(seed)
    su = torch.randn(batch, frames, unit_dim, generator=gen)
    sess = torch.randint(0, num_sessions, (batch,), generator=gen)
    x_real = torch.tanh(torch.randn(batch, frames * hop, channels, generator=gen))
    return su, sess, x_real


def rel_l2(a: Tensor, b: Tensor) -> float:
    """Relative L2 error ||a-b|| / ||b|| in fp64 (the north_star's parity metric)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)
