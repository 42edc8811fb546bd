"""Device time of the captured train step at the bench configuration (quick A/B runs under environment knobs):
   python tools/step_time.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200.synthetic import synthetic_batch
from ste_gan_b200.models.discriminator import DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.trainer import GanTrainer
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8).cuda()
torch.manual_seed(0); d = DiscriminatorSmall(8).cuda()
tr = GanTrainer(g, d, precision="bf16")
batches = [[t.cuda() for t in synthetic_batch(16, 100, seed=s)] for s in range(4)]
tr.capture(16, 100)
for i in range(5):
    tr.step_graph(*batches[i % 4])
torch.cuda.synchronize()
ts = []
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        tr.step_graph(*batches[i % 4])
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 20)
tr.flush(); torch.cuda.synchronize()
knobs = {k: v for k, v in os.environ.items() if k.startswith("STG_")}
print(f"{min(ts):.4f} ms/step (min of {reps} x 20)  {knobs}  loss_g {tr.losses()['loss_g']:.4f}")
