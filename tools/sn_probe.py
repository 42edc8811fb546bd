"""fp32-mode accuracy of the spectral-normed scale discriminator vs the fp64 oracle (GPU):
feature-map error, LeakyReLU sign flips and per-parameter gradient error for ONE forward/backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ste_gan_oracle as O
from ste_gan_b200 import ops, passes
from ste_gan_b200.models.discriminator import DiscriminatorSmall

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dt = torch.float32 if prec == "fp32" else torch.bfloat16
torch.manual_seed(0); d = DiscriminatorSmall(8)
sd = {k: v.detach().clone().double() for k, v in d.state_dict().items()}
_, _, x = O.synthetic_batch(2, 100, seed=3)
# oracle: one training forward on x, loss = sum_i mse(logits_i, 1)
dd = {k: (v if O.is_buffer_key(sd, k) else v.clone().requires_grad_(True)) for k, v in sd.items()}
res_o = O.discriminator_forward(dd, x.double(), True)
loss = sum(((r[-1] - 1) ** 2).mean() for r in res_o)
names = [k for k in dd if not O.is_buffer_key(sd, k)]
grads = dict(zip(names, torch.autograd.grad(loss, [dd[k] for k in names])))
# ours
d = d.cuda()
folds = passes.fold_discriminator(d, dt, training=True)
res, ctx = passes.discriminator_forward(d, x.cuda(), dt, folds)
slots = torch.zeros(8, device="cuda")
dl = []
for fm in res:
    g = torch.empty(fm[-1].shape, device="cuda", dtype=dt)
    ops.mse_const(fm[-1], 1.0, slots[0:1], 1.0, g)
    dl.append(g)
passes.discriminator_backward(d, ctx, dl, None, want_input_grad=False, want_weight_grad=True)
print("loss ours", float(slots[0]), "oracle", float(loss))
subs = passes.disc_subnets(d)
for di, (fm_o, fm_m) in enumerate(zip(res_o, res)):
    kind, sub = subs[di]
    line = []
    for a, b in zip(fm_o, fm_m):
        b = passes.to_reference_layout(b.float().cpu(), kind, getattr(sub, "period", 1)).double()
        flips = int(((a > 0) != (b > 0)).sum())
        line.append(f"{O.rel_l2(b, a):.1e}/{flips}")
    print(f"disc {di} fmap err/flips:", " ".join(line))
for k, p in d.named_parameters():
    if k.startswith("multi_scale_disc.0") or k.endswith("layers.0.weight_v"):
        print(f"  {k:48s} {O.rel_l2(p.grad.cpu(), grads[k]):.3e}")
for k in ("weight_u", "weight_v"):
    for j in range(4):
        n = f"multi_scale_disc.0.layers.{j}.{k}"
        print(f"  {n:48s} {O.rel_l2(d.state_dict()[n].cpu(), dd[n]):.3e}")
