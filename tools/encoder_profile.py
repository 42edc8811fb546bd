"""Kernel breakdown of the EMG-encoder loss pass (forward + input gradient) at the bench shape (B=16, 1600 samples), bf16."""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from ste_gan_b200 import passes_encoder as pe
from ste_gan_b200.models.emg_encoder import EMGEncoderTransformer

torch.manual_seed(0)
enc = EMGEncoderTransformer(8, 256, 48).eval().cuda()
plan = enc.plan(torch.bfloat16)
g = torch.Generator().manual_seed(1)
x = torch.tanh(torch.randn(16, 1600, 8, generator=g)).cuda()
ut = torch.randn(16, 100, 256, generator=g).cuda(); ph = torch.randint(0, 48, (16, 100), generator=g).cuda()
slots = torch.zeros(2, device="cuda")
run = lambda: pe.encoder_losses(plan, x, ut, ph, slots)
for _ in range(3):
    run()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    run()
gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    gr.replay()
e1.record(); torch.cuda.synchronize()
print(f"encoder loss pass (graph replay): {e0.elapsed_time(e1) / 10:.3f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gr.replay(); torch.cuda.synchronize()
path = tempfile.mktemp(suffix=".json"); prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
agg = {}
for e in ev:
    import re
    ids = [w for w in re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*[<(]", e["name"]) if w not in ("void", "anonymous")]
    k = ids[0] if ids else e["name"][:40]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k:36s} {n:4d} launches {us:9.1f} us")
