"""One eager GAN train step at the bench configuration (batch 16, 100 frames, bf16) for ncu:
   python tools/profile_step.py [n_steps]   (step 0 is the warm-up; profile with -s <launches of step 0>)
Prints the library's launch counter after every step so the ncu skip count can be set exactly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200.synthetic import synthetic_batch
from ste_gan_b200 import _lib
from ste_gan_b200.models.discriminator import DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.trainer import GanTrainer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lib = _lib.load(build_if_missing=False)
torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8).cuda()
torch.manual_seed(0); d = DiscriminatorSmall(8).cuda()
tr = GanTrainer(g, d, precision="bf16")
batch = [t.cuda() for t in synthetic_batch(16, 100, seed=0)]
import json
from ste_gan_b200 import ops
for i in range(n):
    if i == n - 1:
        ops.profile = []          # record (kind, engine, shape) of every conv / wgrad call of the last step, in launch order
    tr.step(*batch)
    torch.cuda.synchronize()
    print(f"step {i}: library launches so far {lib.stg_launch_count()}", flush=True)
prof, ops.profile = ops.profile, None
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/step_shapes.json", "w") as f:
    json.dump([dict(kind=p["kind"], engine=p["engine"], flops=p["flops"], bytes=p["bytes"], shape=p["shape"],
                    ms_event=p["events"][0].elapsed_time(p["events"][1])) for p in prof], f)
print(tr.losses())
