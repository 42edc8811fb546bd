"""bf16-vs-fp32 error budget of one train step (GPU).  Prints relative L2 of every feature map,
logit tensor and parameter gradient of the bf16 mode against this library's own fp32 mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ste_gan_oracle as O
from ste_gan_b200 import passes
from ste_gan_b200.models.discriminator import DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.trainer import GanTrainer


def nets():
    torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8)
    torch.manual_seed(0); d = DiscriminatorSmall(8)
    return g.cuda(), d.cuda()


B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
su, sess, x_real = (t.cuda() for t in O.synthetic_batch(B, 100, seed=3))
out = {}
for prec in ("fp32", "bf16"):
    g, d = nets()
    tr = GanTrainer(g, d, precision=prec)
    dt = tr.dtype
    x_pred, _ = passes.generator_forward(g, su, sess, None, dt, False)
    folds = passes.fold_discriminator(d, dt, training=False)
    res, _ = passes.discriminator_forward(d, x_real, dt, folds)
    g2, d2 = nets()
    tr = GanTrainer(g2, d2, precision=prec)
    tr._phase_d(su, sess, None, x_real)
    gd = {n: p.grad.detach().float().clone() for n, p in d2.named_parameters()}
    tr._phase_g(x_real, update_d=False)
    gg = {n: p.grad.detach().float().clone() for n, p in g2.named_parameters()}
    out[prec] = dict(x_pred=x_pred.float(), fmaps=[[f.float() for f in fm] for fm in res], gd=gd, gg=gg, L=tr.losses())
a, b = out["fp32"], out["bf16"]
print("x_pred", O.rel_l2(b["x_pred"], a["x_pred"]))
for i, (fa, fb) in enumerate(zip(a["fmaps"], b["fmaps"])):
    print("disc", i, " ".join(f"{O.rel_l2(y, x):.2e}" for x, y in zip(fa, fb)))
print("losses fp32", a["L"]); print("losses bf16", b["L"])
errs = sorted(((O.rel_l2(b["gd"][k], a["gd"][k]), k) for k in a["gd"]), reverse=True)
print("grad_d worst:"); [print(f"  {e:.3e} {k}  |g|={a['gd'][k].norm():.3e}") for e, k in errs[:25]]
errs = sorted(((O.rel_l2(b["gg"][k], a["gg"][k]), k) for k in a["gg"]), reverse=True)
print("grad_g worst:"); [print(f"  {e:.3e} {k}  |g|={a['gg'][k].norm():.3e}") for e, k in errs[:12]]
