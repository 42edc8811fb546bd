"""One eager GAN train step at the bench configuration (batch 16, 100 frames, bf16) for ncu:
   python tools/profile_step.py [n_steps]   (step 0 is the warm-up; profile with -s <launches of step 0>)
Prints the library's launch counter after every step so the ncu skip count can be set exactly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ste_gan_oracle as O
from ste_gan_b200 import _lib
from ste_gan_b200.models.discriminator import DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.trainer import GanTrainer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lib = _lib.load(build_if_missing=False)
torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8).cuda()
torch.manual_seed(0); d = DiscriminatorSmall(8).cuda()
tr = GanTrainer(g, d, precision="bf16")
batch = [t.cuda() for t in O.synthetic_batch(16, 100, seed=0)]
for i in range(n):
    tr.step(*batch)
    torch.cuda.synchronize()
    print(f"step {i}: library launches so far {lib.stg_launch_count()}", flush=True)
print(tr.losses())
