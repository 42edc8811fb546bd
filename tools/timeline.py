"""Kernel timeline of one CUDA-graph replay of the train step (bench configuration) from torch.profiler / CUPTI:
   python tools/timeline.py [out.json]  -> per-kernel (start us, duration us, stream, name) list + a summary."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from oracle import ste_gan_oracle as O
from ste_gan_b200.models.discriminator import DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.trainer import GanTrainer

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.json"
torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8).cuda()
torch.manual_seed(0); d = DiscriminatorSmall(8).cuda()
tr = GanTrainer(g, d, precision="bf16")
batch = [t.cuda() for t in O.synthetic_batch(16, 100, seed=0)]
tr.capture(16, 100)
for _ in range(5):
    tr.step_graph(*batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        tr.step_graph(*batch)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace.json")
ev = json.load(open("/tmp/trace.json"))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]
rows = [dict(ts=round(e["ts"] - t0, 2), dur=round(e["dur"], 2), stream=e["args"].get("stream"), name=e["name"][:60]) for e in ks]
json.dump(rows, open(out, "w"))
print(len(rows), "gpu events; span", rows[-1]["ts"] + rows[-1]["dur"], "us")
