"""Device time of every captured piece of the pipelined train step, each replayed alone (bench configuration):
a1 D folds | a2 real pass of the spectral-norm stack | b1 G fold + forward | b2 fake pass + batched stacks | b3 D backward +
fold backward | g2 phase G (D AdamW, D passes, losses, D dgrad, G backward, fold backward) | g3 G AdamW.
    python tools/piece_times.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200.synthetic import synthetic_batch
from ste_gan_b200.models.discriminator import DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.trainer import GanTrainer

torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8).cuda()
torch.manual_seed(0); d = DiscriminatorSmall(8).cuda()
tr = GanTrainer(g, d, precision="bf16")
batch = [t.cuda() for t in synthetic_batch(16, 100, seed=0)]
tr.capture(16, 100, pipelined=True)
for _ in range(5):
    tr.step_graph(*batch)
tr.flush(); torch.cuda.synchronize()
a1, a2, b1, b2, b3 = tr._d_graphs
g1, g2, g3 = tr._graphs
pieces = [("a1 D folds", [a1]), ("a2 D real (sn stack)", [a2]), ("b1 G fold+fwd", [b1]), ("b2 D fake + batched", [b2]),
          ("b3 D bwd", [b3]), ("g2 phase G", g2), ("g3 G adamw", [g3])]
n = 20
tot = 0.0
for name, gs in pieces:
    t = 0.0
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for x in gs:
            x.replay()
        e1.record(); torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
    print(f"{name:24s} {1e3 * t / n:8.1f} us")
    tot += t / n
print(f"sum {tot:.3f} ms (a1, a2 and g3 overlap other pieces in the pipelined step)")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    tr.step_graph(*batch)
tr.flush(); e1.record(); torch.cuda.synchronize()
print(f"pipelined step_graph {e0.elapsed_time(e1) / n:.3f} ms per step")
tr.capture(16, 100)          # the default: one fused graph per step
for _ in range(5):
    tr.step_graph(*batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    tr.step_graph(*batch)
tr.flush(); e1.record(); torch.cuda.synchronize()
print(f"fused step_graph {e0.elapsed_time(e1) / n:.3f} ms per step")
