"""Synthetic inputs of the benchmark / smoke workloads: soft speech units ~ N(0,1) as in the reference's
own smoke block (models/discriminator.py:217), session ids uniform over the 17 sessions, real EMG tanh-ranged
(configs/data/gaddy_and_klein_corpus.yaml:5).  Same generator sequence as oracle.ste_gan_oracle.synthetic_batch
(tests/test_host_cpu.py pins the two against each other) - kept here so that the GPU arm of bench.py does not import
the oracle."""
import torch


def synthetic_batch(batch: int, frames: int, seed: int = 0, channels: int = 8, unit_dim: int = 256, hop: int = 16,
                    num_sessions: int = 17):
    gen = torch.Generator().manual_seed(seed)
    su = torch.randn(batch, frames, unit_dim, generator=gen)
    sess = torch.randint(0, num_sessions, (batch,), generator=gen)
    x_real = torch.tanh(torch.randn(batch, frames * hop, channels, generator=gen))
    return su, sess, x_real
