"""Reference arms of bench.py: the reference's OWN modules (or, where they are not installed, the oracle port)
driven through the train step of ste_gan/train.py:165-268, on the host cores or - for the `context.torch_gpu`
block - on the GPU through the container's torch + cuDNN.  Measurement infrastructure only: nothing under
ste_gan_b200/ imports this file.

kind "reference": the UNMODIFIED reference package, installed once with
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /tmp/<copy of /root/reference>
(baseline/_ref is git-ignored but travels to the GPU box).  The reference imports `omegaconf` for type annotations and
`"params" in cfg.model` only (models/generator.py:7,190; models/discriminator.py:8); omegaconf is not in the offline
wheelhouse, so a 3-line in-memory stub stands in for it and the factories get an attribute-dict cfg.
`ste_gan/train.py` itself cannot be imported (omegaconf + matplotlib + tensorboard + a dataset on disk), so its inner
loop body is restated here around the reference's modules: netG / netD from init_emg_generator / init_emg_discriminators,
MultiTimeDomainFeatureLoss, two torch.optim.AdamW(lr 2e-4, betas (.8,.99)) (constants.py:57), autocast + GradScaler as
in train.py:151,181,204.
kind "port": oracle/ste_gan_oracle.py (a functional restatement pinned to the reference by tests/golden/).
"""
from __future__ import annotations

import os
import sys
import time
import types

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W_TD, W_FM = 15.0, 7.0          # configs/ste_gan_base_gantts.yaml:33,37


class _Cfg(dict):
    """attribute access + `in` (what the factories use of a DictConfig)"""
    __getattr__ = dict.__getitem__


def load_reference():
    """Import the reference package from baseline/_ref (or /root/reference in the authoring container).  Returns a
    namespace of what the step needs, or None when it is not there."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "ste_gan")):
            break
    else:
        return None
    if "omegaconf" not in sys.modules:
        m = types.ModuleType("omegaconf"); m.OmegaConf = object; m.DictConfig = dict
        sys.modules["omegaconf"] = m
    if cand not in sys.path:
        sys.path.insert(0, cand)
    try:
        from ste_gan.losses.time_domain_loss import MultiTimeDomainFeatureLoss
        from ste_gan.models.discriminator import init_emg_discriminators
        from ste_gan.models.generator import init_emg_generator
    except Exception:      # noqa: BLE001 - a missing transitive dependency means "not available", not a crash
        return None
    return types.SimpleNamespace(init_g=init_emg_generator, init_d=init_emg_discriminators, mtd=MultiTimeDomainFeatureLoss,
                                 where=cand)


def base_cfg(small: bool = True) -> _Cfg:
    """configs/ste_gan_base_gantts.yaml (model section) + configs/data/gaddy_and_klein_corpus.yaml (channel / session counts)"""
    return _Cfg(model=_Cfg(type="EMGGeneratorGanTTS", speech_feature_type="SPEECH_UNITS", discriminator_small=small),
                data=_Cfg(num_emg_channels=8, num_emg_sessions=17))


class RefTrainer:
    """train.py:165-268 around the reference modules.  amp: None (fp32), torch.bfloat16 or torch.float16 (the reference's
    own mixed_precision mode: fp16 autocast + GradScaler)."""
    kind = "reference"

    def __init__(self, ref, device="cpu", small: bool = True, amp=None, compile_: bool = False):
        self.device = torch.device(device)
        cfg = base_cfg(small)
        torch.manual_seed(0); self.netG = ref.init_g(cfg).to(self.device)          # train.py:54
        torch.manual_seed(0); self.netD = ref.init_d(cfg).to(self.device)          # train.py:55
        self.mtd = ref.mtd(8).to(self.device)                                      # train.py:63
        self.optG = torch.optim.AdamW(self.netG.parameters(), lr=2e-4, betas=(0.8, 0.99))   # train.py:80, constants.py:57
        self.optD = torch.optim.AdamW(self.netD.parameters(), lr=2e-4, betas=(0.8, 0.99))   # train.py:81
        self.amp = amp
        self.scaler = torch.amp.GradScaler(self.device.type, enabled=amp == torch.float16)  # train.py:151
        self.fwdG, self.fwdD = self.netG, self.netD
        if compile_:                                                                # train.py:140-146
            self.fwdG, self.fwdD = torch.compile(self.netG), torch.compile(self.netD)
        if self.device.type == "cuda":
            torch.backends.cudnn.benchmark = True                                   # train.py:135

    def _ac(self):
        return torch.autocast(device_type=self.device.type, dtype=self.amp or torch.bfloat16, enabled=self.amp is not None)

    def _d_phase(self, x_pred, x_real):
        with self._ac():
            D_fake_det, D_real = self.fwdD(x_pred.detach()), self.fwdD(x_real)      # train.py:190-191
            loss_D = 0
            for scale in D_fake_det:
                loss_D = loss_D + F.mse_loss(scale[-1], torch.zeros_like(scale[-1]))
            for scale in D_real:
                loss_D = loss_D + F.mse_loss(scale[-1], torch.ones_like(scale[-1]))
            self.scaler.scale(loss_D).backward()                                    # train.py:198-199
            self.scaler.step(self.optD)
        return loss_D

    def _g_losses(self, x_pred, x_real):
        with self._ac():
            D_fake, D_real = self.fwdD(x_pred), self.fwdD(x_real)                   # train.py:206-207
            adv = 0
            for scale in D_fake:
                adv = adv + F.mse_loss(scale[-1], torch.ones_like(scale[-1]))       # :210-211
            td = self.mtd(x_real, x_pred)                                           # :215
            fm = 0
            for i in range(len(D_fake)):
                for j in range(len(D_fake[i]) - 1):
                    fm = fm + F.l1_loss(D_fake[i][j], D_real[i][j].detach())        # :259-262
            loss_G = adv + W_TD * td + W_FM * fm
        return loss_G, adv, td, fm

    def step(self, su, sess, x_real, mode=None):
        dev = self.device
        su, sess, x_real = su.to(dev), sess.to(dev), x_real.to(dev)
        mode = torch.zeros_like(sess) if mode is None else mode.to(dev)
        self.netD.zero_grad(); self.netG.zero_grad()                                # train.py:166-167
        self.netG.train()
        with self._ac():
            x_pred = self.fwdG(su, sess, mode)                                      # train.py:182
        loss_D = self._d_phase(x_pred, x_real)
        loss_G, adv, td, fm = self._g_losses(x_pred, x_real)
        self.scaler.scale(loss_G).backward()                                        # train.py:266-268
        self.scaler.step(self.optG)
        self.scaler.update()
        return dict(loss_d=loss_D, loss_g=loss_G, loss_adv=adv, loss_td=td, loss_fm=fm)

    def disc_losses_step(self, x_pred, x_real):
        """BASELINE.json configs[4]: everything of the step except the generator (see GanTrainer.disc_losses_step)."""
        dev = self.device
        x_pred = x_pred.to(dev).detach().requires_grad_(True)
        x_real = x_real.to(dev)
        self.netD.zero_grad()
        self._d_phase(x_pred, x_real)
        loss_G, *_ = self._g_losses(x_pred, x_real)
        self.scaler.scale(loss_G).backward()
        self.scaler.update()
        return x_pred.grad

    def load_generator_state(self, sd):
        self.netG.load_state_dict({k: v.detach().to(self.device) for k, v in sd.items()})

    @torch.inference_mode()
    def generate(self, su, sess):
        dev = self.device
        with self._ac():
            return self.netG.generate(su.to(dev), sess.to(dev), torch.zeros_like(sess).to(dev))   # generator.py:48-50


class PortTrainer:
    """The oracle's restatement of the same step (kind "port"): used where the reference package is not installed."""
    kind = "port"

    def __init__(self, device="cpu", small: bool = True, amp=None, compile_: bool = False):
        from oracle import ste_gan_oracle as O
        from ste_gan_b200.models.discriminator import Discriminator, DiscriminatorSmall
        from ste_gan_b200.models.generator import EMGGeneratorGanTTS
        self.O, self.device, self.amp, self.small = O, torch.device(device), amp, small
        torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8)
        torch.manual_seed(0); d = (DiscriminatorSmall if small else Discriminator)(8)
        mv = lambda sd: {k: v.detach().clone().to(self.device) for k, v in sd.items()}
        self.ot = O.OracleTrainer(mv(g.state_dict()), mv(d.state_dict()), small=small)
        if self.device.type == "cuda":
            torch.backends.cudnn.benchmark = True

    def _ac(self):
        return torch.autocast(device_type=self.device.type, dtype=self.amp or torch.bfloat16, enabled=self.amp is not None)

    def step(self, su, sess, x_real, mode=None):
        dev = self.device
        with self._ac():
            return self.ot.step(su.to(dev), sess.to(dev), x_real.to(dev))

    def disc_losses_step(self, x_pred, x_real):
        O, ot, dev = self.O, self.ot, self.device
        x_pred = x_pred.to(dev).detach().requires_grad_(True)
        x_real = x_real.to(dev)
        ot.opt_d.zero_grad()
        with self._ac():
            loss_d = O.lsgan_d_loss(O.discriminator_forward(ot.d, x_pred.detach(), self.small),
                                    O.discriminator_forward(ot.d, x_real, self.small))
            loss_d.backward(); ot.opt_d.step()
            d_fake, d_real = O.discriminator_forward(ot.d, x_pred, self.small), O.discriminator_forward(ot.d, x_real, self.small)
            loss_g = O.lsgan_g_loss(d_fake) + W_TD * O.multi_td_loss(x_real, x_pred)[0] + W_FM * O.feature_matching_loss(d_fake, d_real)
        loss_g.backward()
        return x_pred.grad

    def load_generator_state(self, sd):
        self.ot.g = {k: v.detach().clone().to(self.device) for k, v in sd.items()}

    @torch.inference_mode()
    def generate(self, su, sess):
        dev = self.device
        with self._ac():
            return self.O.generator_forward(self.ot.g, su.to(dev), sess.to(dev))


def make_trainer(device="cpu", small: bool = True, amp=None, compile_: bool = False, prefer_reference: bool = True):
    ref = load_reference() if prefer_reference else None
    if ref is not None:
        return RefTrainer(ref, device, small, amp, compile_)
    return PortTrainer(device, small, amp, compile_)


def losses_to_float(d: dict) -> dict:
    return {k: float(v) for k, v in d.items()}


def time_steps(fn, n: int, device) -> float:
    """mean seconds per call of fn() over n calls (wall clock, synchronised on CUDA)"""
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
