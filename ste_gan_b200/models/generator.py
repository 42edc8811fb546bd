"""EMG generator - drop-in mirror of ste_gan/models/generator.py.

Same classes (`EMGGenerator`, `EMGGeneratorGanTTS`), factory (`init_emg_generator`),
constructor arguments, attributes, parameter names / shapes / registration order and random
initialisation as the reference (generator.py:19-195); `forward` runs the fused CUDA passes
(ste_gan_b200/passes.py) instead of 45 cuDNN convolutions + ~60 elementwise kernels.
"""
from typing import Dict

import torch
import torch.nn as nn
from torch import Tensor

import ste_gan_b200 as ste_gan
from ste_gan_b200.constants import DataType
from ste_gan_b200.layers.conv import GBlock, WNConv1d


class EMGGenerator(nn.Module):
    """Base class for the EMG generator (generator.py:19-75)."""

    def __init__(self, speech_feature_type: str, speech_input_dim: int, num_sessions: int,
                 num_output_channels: int) -> None:
        super().__init__()
        self.speech_feature_type = speech_feature_type
        self.speech_input_dim = speech_input_dim
        self.num_output_channels = num_output_channels
        self.num_sessions = num_sessions

    def forward(self, speech_unit_sequence: Tensor, session_ids: Tensor, speaking_mode_ids: Tensor, ar=None):
        raise NotImplementedError("Must implement the EMGGenerator interface.")

    @torch.inference_mode()
    def generate(self, speech_unit_sequence: Tensor, session_ids: Tensor, speaking_mode_ids: Tensor) -> Tensor:
        return self(speech_unit_sequence, session_ids, speaking_mode_ids)

    @torch.inference_mode()
    def generate_from_data_dict(self, data_dict: Dict, device: torch.device) -> Tensor:
        """generator.py:52-75: EMGDataset item -> generated EMG [T, C] on the CPU."""
        s_t = data_dict[self.speech_feature_type].to(device)
        sess_idx = data_dict[DataType.SESSION_INDEX].to(device)
        spk_mode_idx = data_dict[DataType.SPEAKING_MODE_INDEX].to(device)
        if len(s_t.shape) == 2:
            s_t = s_t.unsqueeze(0)
            sess_idx = sess_idx.unsqueeze(0)
            spk_mode_idx = spk_mode_idx.unsqueeze(0)
        return self.generate(s_t, sess_idx, spk_mode_idx).squeeze(0).detach().cpu()


class EMGGeneratorGanTTS(EMGGenerator):
    """The GAN-TTS / CARGAN style generator of the paper (generator.py:78-162)."""

    def __init__(self, speech_feature_type: str, speech_input_dim: int, num_sessions: int, num_emg_channels: int,
                 use_speaking_mode_embedding: bool = False, use_session_embeddings: bool = True,
                 num_speaking_modes: int = 3, embedding_dim: int = 64, channels: int = 768):
        super().__init__(speech_feature_type=speech_feature_type, speech_input_dim=speech_input_dim,
                         num_sessions=num_sessions, num_output_channels=num_emg_channels)
        self.session_embeddings = nn.Embedding(num_sessions, embedding_dim) if use_session_embeddings else None
        self.use_session_embeddings = use_session_embeddings
        self.speaking_mode_embeddings = (nn.Embedding(num_speaking_modes, embedding_dim)
                                         if use_speaking_mode_embedding else None)
        self.use_speaking_mode_embedding = use_speaking_mode_embedding
        self.input_size = self.speech_input_dim + ((use_session_embeddings * embedding_dim)
                                                   + (use_speaking_mode_embedding * embedding_dim))
        upsample_last = 2 if self.speech_feature_type == DataType.SPEECH_UNITS else 1     # generator.py:116
        self.channels = channels
        self.gblocks = nn.Sequential(
            WNConv1d(self.input_size, channels, kernel_size=1),
            GBlock(channels, channels),
            GBlock(channels, channels),
            GBlock(channels, channels // 2, upsample=2),
            GBlock(channels // 2, channels // 2, upsample=2),
            GBlock(channels // 2, channels // 2, upsample=2),
            GBlock(channels // 2, channels // 4, upsample=upsample_last),
            GBlock(channels // 4, channels // 4),
            GBlock(channels // 4, channels // 4),
        )
        self.last_conv = nn.Sequential(nn.ReLU(), WNConv1d(channels // 4, num_emg_channels, kernel_size=3, padding=1))

    def forward(self, speech_unit_sequence: Tensor, session_ids: Tensor, speaking_mode_ids: Tensor):
        """[B,T,D] units, [B] session ids, [B] speaking-mode ids -> [B, 16T (8T for MFCCs), C] in (-1, 1)."""
        from ste_gan_b200.autograd import GeneratorFn
        return GeneratorFn.apply(self, speech_unit_sequence, session_ids, speaking_mode_ids, *self.parameters())


def init_emg_generator(cfg, emg_generator_type: str = "") -> EMGGenerator:
    """Factory (generator.py:165-195); cfg is the reference's DictConfig (attribute access + `in`)."""
    speech_feature_type = cfg.model.speech_feature_type
    if speech_feature_type == DataType.SPEECH_UNITS:
        speech_input_dim: int = ste_gan.SPEECH_UNITS_FEAT_SIZE
    elif speech_feature_type == DataType.MFCCS:
        speech_input_dim: int = ste_gan.NUM_MFCCS
    else:
        raise ValueError(f"Unrecognized speech feature type: {speech_feature_type}")
    num_emg_channels = cfg.data.num_emg_channels
    num_sessions = cfg.data.num_emg_sessions
    if not emg_generator_type:
        emg_generator_type = cfg.model.type
    params = dict(num_emg_channels=num_emg_channels, num_sessions=num_sessions,
                  speech_feature_type=speech_feature_type, speech_input_dim=speech_input_dim)
    extra_params = cfg.model.params if "params" in cfg.model else {}
    if emg_generator_type == "EMGGeneratorGanTTS":
        return EMGGeneratorGanTTS(**params, **extra_params)
    else:
        raise ValueError(f"Unrecognized EMG generator type: {emg_generator_type}")
