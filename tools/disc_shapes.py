"""Per-shape conv / wgrad times of one eager single-stream GanTrainer.disc_losses_step (configs[4]).  python tools/disc_shapes.py [small|full] [batch]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200 import ops
from ste_gan_b200.models.discriminator import Discriminator, DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.synthetic import synthetic_batch
from ste_gan_b200.trainer import GanTrainer
family = sys.argv[1] if len(sys.argv) > 1 else "full"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8, channels=64).cuda()
torch.manual_seed(0); d = (DiscriminatorSmall if family == "small" else Discriminator)(8).cuda()
tr = GanTrainer(g, d, precision="bf16")
tr.concurrent_d = False
x_real = synthetic_batch(B, 100, seed=11)[2].cuda()
x_pred = torch.tanh(torch.randn(B, 1600, 8, generator=torch.Generator().manual_seed(12))).cuda()
for _ in range(2):
    tr.disc_losses_step(x_pred, x_real)
torch.cuda.synchronize()
ops.profile = []
torch.cuda._sleep(int(0.06 * 1.9e9))
tr.disc_losses_step(x_pred, x_real)
torch.cuda.synchronize()
prof, ops.profile = ops.profile, None
agg = collections.OrderedDict()
for p in prof:
    k = (p["kind"], p["engine"], p["shape"])
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += p["events"][0].elapsed_time(p["events"][1]) * 1e3; a[2] += p["flops"]
tot = sum(a[1] for a in agg.values())
print(f"{family} B={B}: {len(prof)} conv/wgrad launches, {tot:.0f} us (alone, single stream)")
print("kind  engine  (B, phases, t_src, t_dst, c_src, c_dst, k, dil, stride, groups)   n   total_us  avg_us  TFLOP/s")
for (kind, eng, shape), (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{kind:5s} {eng:7s} {str(shape):58s} {n:3d} {us:9.1f} {us / n:7.1f} {fl / us / 1e6:8.1f}")
