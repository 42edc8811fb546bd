"""Kernel timeline of one CUDA-graph replay of the train step (bench configuration) from torch.profiler / CUPTI:
   python tools/timeline.py [out.json]  -> per-kernel (start us, duration us, stream, name) list + a summary."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from ste_gan_b200.synthetic import synthetic_batch
from ste_gan_b200.models.discriminator import DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.trainer import GanTrainer

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.json"
rank = 0
if int(os.environ.get("WORLD_SIZE", "1")) > 1:      # under torchrun: the data-parallel step, rank 0 writes its timeline
    from ste_gan_b200.dist import init_from_env
    rank, world, local = init_from_env("nccl")
    torch.cuda.set_device(local)
torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8).cuda()
torch.manual_seed(0); d = DiscriminatorSmall(8).cuda()
tr = GanTrainer(g, d, precision="bf16")
tr.reducer.broadcast(tr.G.flat); tr.reducer.broadcast(tr.D.flat)
batch = [t.cuda() for t in synthetic_batch(16, 100, seed=rank)]
tr.capture(16, 100)
for _ in range(5):
    tr.step_graph(*batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        tr.step_graph(*batch)
    torch.cuda.synchronize()
if rank != 0:
    sys.exit(0)
prof.export_chrome_trace("/tmp/trace.json")
ev = json.load(open("/tmp/trace.json"))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]
rows = [dict(ts=round(e["ts"] - t0, 2), dur=round(e["dur"], 2), stream=e["args"].get("stream"), name=e["name"][:60],
             grid=e["args"].get("grid"), block=e["args"].get("block"), smem=e["args"].get("shared memory"),
             regs=e["args"].get("registers per thread")) for e in ks]
json.dump(rows, open(out, "w"))
print(len(rows), "gpu events; span", rows[-1]["ts"] + rows[-1]["dur"], "us")
