"""Where does the bf16 error of the first-layer discriminator gradients come from?  (GPU)
Records, for every discriminator conv in the D phase, the wgrad operands (x, dy) and the packed dw in fp32 and
bf16 mode and prints relative L2 of each: dy error, dw error (pre-fold), final parameter-gradient error, and
the fraction of dw that lies along the weight direction (the part the weight-norm fold projects away)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ste_gan_oracle as O
from ste_gan_b200 import passes
from ste_gan_b200.models.discriminator import DiscriminatorSmall
from ste_gan_b200.models.generator import EMGGeneratorGanTTS
from ste_gan_b200.trainer import GanTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
su, sess, x_real = (t.cuda() for t in O.synthetic_batch(B, 100, seed=3))
rec = {}
orig_wgrad = passes._wgrad


def run(prec):
    torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8).cuda()
    torch.manual_seed(0); d = DiscriminatorSmall(8).cuda()
    names = {id(m): n for n, m in d.named_modules()}
    out = {}
    calls = {}

    def wg(f, x, dy, B_, t_x, t_dy, ws, phases=1):
        off0 = ws.off
        orig_wgrad(f, x, dy, B_, t_x, t_dy, ws, phases=phases)
        m = f.mod
        n = m.out_channels * m.kernel * (m.in_channels // m.groups)
        nm = names.get(id(m), "?")
        c = calls.get(nm, 0); calls[nm] = c + 1
        out[(nm, c)] = dict(x=x.float().clone(), dy=dy.float().clone(), dw=ws.buf[off0:off0 + n].clone())
    passes._wgrad = wg
    tr = GanTrainer(g, d, precision=prec)
    tr._phase_d(su, sess, None, x_real)
    passes._wgrad = orig_wgrad
    gd = {n: p.grad.detach().float().clone() for n, p in d.named_parameters()}
    return out, gd, d


a, ga, d = run("fp32")
b, gb, _ = run("bf16")
mods = dict(d.named_modules())
print(f"{'layer':44s} pass  err(x)    err(dy)   err(dw)   radial_frac")
for key in a:
    nm, c = key
    if not (nm.endswith("layers.0") or nm.endswith("layers.1") or nm.endswith("output")):
        continue
    m = mods[nm]
    ex, edy, edw = O.rel_l2(b[key]["x"], a[key]["x"]), O.rel_l2(b[key]["dy"], a[key]["dy"]), O.rel_l2(b[key]["dw"], a[key]["dw"])
    # radial fraction: |proj of dw on v| / |dw| per output channel (weight-norm only)
    rf = float("nan")
    if m.norm == "weight_norm":
        v = m.weight_v.data.view(m.out_channels, m.in_channels // m.groups, m.kernel).permute(0, 2, 1).reshape(m.out_channels, -1)
        dw = a[key]["dw"].view(m.out_channels, -1)
        vh = v / v.norm(dim=1, keepdim=True)
        rad = (dw * vh).sum(1, keepdim=True) * vh
        rf = float(rad.norm() / dw.norm())
    print(f"{nm:44s} {c}    {ex:.2e}  {edy:.2e}  {edw:.2e}  {rf:.3f}")
print("final param grads (bf16 vs fp32):")
for k in ga:
    if ".layers.0." in k:
        print(f"  {k:50s} {O.rel_l2(gb[k], ga[k]):.3e}")
