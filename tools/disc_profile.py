"""Kernel breakdown of GanTrainer.disc_losses_step (BASELINE.json configs[4]) for one discriminator family and batch size, bf16.
    python tools/disc_profile.py [small|full] [batch]"""
import os, sys, json, re, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from ste_gan_b200.models.discriminator import Discriminator, DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.synthetic import synthetic_batch
from ste_gan_b200.trainer import GanTrainer

family = sys.argv[1] if len(sys.argv) > 1 else "full"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8, channels=64).cuda()
torch.manual_seed(0); d = (DiscriminatorSmall if family == "small" else Discriminator)(8).cuda()
tr = GanTrainer(g, d, precision="bf16")
x_real = synthetic_batch(B, 100, seed=11)[2].cuda()
x_pred = torch.tanh(torch.randn(B, 1600, 8, generator=torch.Generator().manual_seed(12))).cuda()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        tr.disc_losses_step(x_pred, x_real)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    tr.disc_losses_step(x_pred, x_real)
for _ in range(3):
    gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    gr.replay()
e1.record(); torch.cuda.synchronize()
print(f"{family} D + losses, batch {B}: {e0.elapsed_time(e1) / 10:.3f} ms per step")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gr.replay(); torch.cuda.synchronize()
path = tempfile.mktemp(suffix=".json"); prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
agg = {}
for e in ev:
    ids = [w for w in re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*[<(]", e["name"]) if w not in ("void", "anonymous")]
    k = ids[0] if ids else e["name"][:40]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
span = max(e["ts"] + e["dur"] for e in ev) - min(e["ts"] for e in ev)
print(f"span {span:.0f} us, busy {sum(a[1] for a in agg.values()):.0f} us, {len(ev)} kernels")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{k:36s} {n:4d} launches {us:9.1f} us")
