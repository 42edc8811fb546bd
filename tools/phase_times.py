"""Device time of the three CUDA graphs of one train step (phase D, phase G, G optimiser) at the bench configuration,
with and without the multi-stream schedule.  python tools/phase_times.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200 import synthetic as O
from ste_gan_b200.models.discriminator import DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.trainer import GanTrainer

for conc in (True, False):
    torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8).cuda()
    torch.manual_seed(0); d = DiscriminatorSmall(8).cuda()
    tr = GanTrainer(g, d, precision="bf16")
    tr.concurrent_d = conc
    batch = [t.cuda() for t in O.synthetic_batch(16, 100, seed=0)]
    tr.capture(16, 100)
    for _ in range(5):
        tr.step_graph(*batch)
    torch.cuda.synchronize()
    tr.flush(); torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tr.step_graph(*batch)
    tr.flush(); e1.record(); torch.cuda.synchronize()
    print(f"concurrent={conc}: pipelined step_graph {e0.elapsed_time(e1)/n:.3f} ms per step")
    tr.capture(16, 100, pipelined=False)
    for _ in range(3):
        tr.step_graph(*batch)
    torch.cuda.synchronize()
    tot = [0.0, 0.0, 0.0]
    for _ in range(n):
        for i, gr in enumerate(tr._graphs):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for g_ in (gr if isinstance(gr, list) else [gr]):    # phase G: one graph per gradient bucket
                g_.replay()
            e1.record(); torch.cuda.synchronize()
            tot[i] += e0.elapsed_time(e1)
    print(f"concurrent={conc}: phase D {tot[0]/n:.3f} ms, phase G {tot[1]/n:.3f} ms, G optimiser {tot[2]/n:.3f} ms, sum {sum(tot)/n:.3f} ms")
