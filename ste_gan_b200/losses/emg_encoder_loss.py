"""Losses through the frozen EMG encoder - drop-in mirror of ste_gan/losses/emg_encoder_loss.py.

`EMGEncoderLoss(emg_encoder)(emg, speech_unit_target, phoneme_target) -> EMGEncoderLossOutput` with the reference's
fields and counting properties (emg_encoder_loss.py:19-52); the encoder forward, the two losses and the gradient w.r.t. the
EMG signal run in the CUDA library (ste_gan_b200/passes_encoder.py, csrc/encoder.cu).  The fused train step
(`GanTrainer(emg_encoder=...)`) calls the same pass directly, without autograd.
"""
from dataclasses import dataclass

import torch
import torch.nn as nn
from torch import Tensor

from ste_gan_b200.models.emg_encoder import EMGEncoder, SILENCE_PHONEME_INDEX


@dataclass
class EMGEncoderLossOutput:
    speech_unit_pred: Tensor
    phoneme_pred: Tensor
    speech_unit_loss: Tensor
    phoneme_loss: Tensor
    phoneme_targets: Tensor

    @property
    def num_phones(self) -> int:
        return len(torch.flatten(self.phoneme_targets))

    @property
    def num_silence_phones(self) -> int:
        return (torch.flatten(self.phoneme_targets) == SILENCE_PHONEME_INDEX).sum().item()

    @property
    def num_correct_phones(self) -> int:
        pred = torch.flatten(self.phoneme_pred.argmax(-1))
        return int((pred == torch.flatten(self.phoneme_targets)).sum().item())

    @property
    def num_correct_phones_no_silence(self) -> int:
        pred, target = torch.flatten(self.phoneme_pred.argmax(-1)), torch.flatten(self.phoneme_targets)
        return ((pred == target) & (target != SILENCE_PHONEME_INDEX)).sum().item()


class EMGEncoderLoss(nn.Module):
    """emg_encoder_loss.py:56-84."""

    def __init__(self, emg_encoder: EMGEncoder) -> None:
        super().__init__()
        self.emg_encoder = emg_encoder
        self.emg_encoder.eval()

    def speech_unit_loss(self, speech_unit_target: Tensor, speech_unit_pred: Tensor) -> Tensor:
        """mean over (b t) rows of ||target - pred + 1e-6||_2 (F.pairwise_distance, emg_encoder_loss.py:63-67)."""
        from ste_gan_b200.autograd import EncoderLossFn
        dummy_logits = torch.zeros(speech_unit_pred.shape[:-1] + (1,), device=speech_unit_pred.device)
        dummy_ph = torch.zeros(speech_unit_pred.shape[:-1], device=speech_unit_pred.device, dtype=torch.int64)
        return EncoderLossFn.apply(speech_unit_pred, speech_unit_target, dummy_logits, dummy_ph)[0]

    def forward(self, emg_signal: Tensor, target_speech_units: Tensor, target_phoneme_sequence: Tensor) -> EMGEncoderLossOutput:
        from ste_gan_b200.autograd import EncoderLossFn
        speech_unit_pred, phoneme_pred = self.emg_encoder(emg_signal)
        su_loss, ph_loss = EncoderLossFn.apply(speech_unit_pred, target_speech_units, phoneme_pred, target_phoneme_sequence)
        return EMGEncoderLossOutput(speech_unit_pred, phoneme_pred, su_loss, ph_loss, target_phoneme_sequence)
