"""Constants of the hot path, same names and values as the reference's ste_gan/constants.py
(:13-57, :192-239) so that `import ste_gan_b200 as ste_gan` style call sites keep working."""
import functools

import torch

EMG_SAMPLE_RATE = 800                 # constants.py:11
BATCH_SIZE = 32                       # constants.py:14
CHUNK_SIZE = 2048                     # constants.py:17
SPEECH_UNITS_FEAT_SIZE = 256          # constants.py:26
NUM_MFCCS = 25                        # constants.py:29
EMBEDDING_DIM_SIZE = 64               # constants.py:32
NUM_EMG_CHANNELS = 8                  # constants.py:35
NUM_EMG_SESSIONS = 17                 # constants.py:38
SPEECH_UNIT_HOPSIZE_SECONDS = 0.02    # constants.py:42
HOPSIZE = int(EMG_SAMPLE_RATE * SPEECH_UNIT_HOPSIZE_SECONDS)  # constants.py:45
OPTIMIZER = functools.partial(torch.optim.AdamW, lr=2e-4, betas=(.8, .99))  # constants.py:57
RANDOM_SEED = 0                       # constants.py:60
LOSS_FEAT_MATCH_WEIGHT = 7.           # constants.py:80


class DataType:
    """Keys of the data dictionaries (constants.py:192-239); only the ones the hot path reads."""
    REAL_EMG = "REAL_EMG"
    SPEECH_UNITS = "SPEECH_UNITS"
    MFCCS = "MFCCS"
    SESSION_INDEX = "SESSION_INDEX"
    SPEAKING_MODE_INDEX = "SPEAKING_MODE_IDX"
    FAKE_EMG = "FAKE_EMG"
    PHONEMES = "PHONEMES"
