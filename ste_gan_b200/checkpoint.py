"""Checkpoint interchange with the reference.

The reference writes three files per save (ste_gan/train.py:421-436, :441-456 for the final one):

    netG-{steps:08d}.pt        netG.state_dict()
    netD-{steps:08d}.pt        netD.state_dict()
    checkpoint-{steps:08d}.pt  {'epoch', 'steps', 'optG': optG.state_dict(), 'optD': optD.state_dict()}

and resumes from the highest-numbered triple (ste_gan/utils/common.py:23-61), stripping the `_orig_mod.` prefix that
torch.compile adds to keys (common.py:13-21).  The drop-in modules already have the reference's state_dict keys, so the
model files are interchangeable as they are.  The optimiser state is not: the reference uses two torch.optim.AdamW
(per-parameter 'step' / 'exp_avg' / 'exp_avg_sq', parameters identified by registration index), the fused trainer keeps
one flat first / second moment buffer and one step counter per network.  The functions here convert between the two, so
a run can move between the reference loop and `GanTrainer` in either direction.

Host-side only (plain tensor copies through torch); nothing here is on the hot path.
"""
from __future__ import annotations

import re
from collections import OrderedDict
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple

import torch

ADAMW_DEFAULTS = dict(lr=2e-4, betas=(0.8, 0.99), eps=1e-8, weight_decay=1e-2, amsgrad=False, foreach=None, maximize=False,
                      capturable=False, differentiable=False, fused=None)      # constants.py:57 + torch.optim.AdamW defaults


def fix_state_dict(state_dict):
    """common.py:13-21: drop the `_orig_mod.` prefix torch.compile leaves in parameter names."""
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k.replace("_orig_mod.", "")] = v
    return out


def flat_to_adamw_state(shapes: List[Tuple[str, torch.Size]], offsets: Dict[str, int], m: torch.Tensor, v: torch.Tensor,
                        step: int, hyper: Optional[dict] = None) -> dict:
    """Flat moments -> the state_dict of a torch.optim.AdamW over the same parameters in registration order.
    shapes: [(name, shape)] in registration order; offsets: name -> offset into the flat buffers."""
    state = {}
    for i, (nm, shp) in enumerate(shapes):
        n, o = int(torch.Size(shp).numel()), offsets[nm]
        state[i] = dict(step=torch.tensor(float(step)), exp_avg=m[o:o + n].detach().reshape(shp).clone().cpu(),
                        exp_avg_sq=v[o:o + n].detach().reshape(shp).clone().cpu())
    group = dict(ADAMW_DEFAULTS)
    group.update(hyper or {})
    group["params"] = list(range(len(shapes)))
    return dict(state=state if step > 0 else {}, param_groups=[group])


def adamw_state_to_flat(sd: dict, shapes: List[Tuple[str, torch.Size]], offsets: Dict[str, int], m: torch.Tensor,
                        v: torch.Tensor) -> int:
    """The inverse: fills the flat moment buffers in place from a torch.optim.AdamW state_dict and returns the step
    count.  Parameters without state (never stepped) get zero moments; all stepped parameters must agree on 'step'
    (the reference steps every parameter of a network together)."""
    m.zero_(); v.zero_()
    ids = sd["param_groups"][0]["params"] if len(sd["param_groups"]) == 1 else [i for g in sd["param_groups"] for i in g["params"]]
    if len(ids) != len(shapes):
        raise ValueError(f"optimizer state has {len(ids)} parameters, the network {len(shapes)}")
    steps = set()
    for pos, (nm, shp) in enumerate(shapes):
        st = sd["state"].get(ids[pos])
        if st is None:
            continue
        if tuple(st["exp_avg"].shape) != tuple(shp):
            raise ValueError(f"{nm}: moment shape {tuple(st['exp_avg'].shape)} != parameter shape {tuple(shp)}")
        n, o = int(torch.Size(shp).numel()), offsets[nm]
        m[o:o + n].copy_(st["exp_avg"].reshape(-1))
        v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
        steps.add(int(float(st["step"])))
    if len(steps) > 1:
        raise ValueError(f"parameters disagree on the step count: {sorted(steps)}")
    return steps.pop() if steps else 0


def latest_step(directory: Path) -> Optional[str]:
    """common.py:31-39: the highest integer suffix among checkpoint-*.pt, as the zero-padded string used in file names."""
    nums = []
    for f in Path(directory).glob("checkpoint-*.pt"):
        mt = re.fullmatch(r"checkpoint-(\d+)", f.stem)
        if mt:
            nums.append(int(mt.group(1)))
    return f"{max(nums):08d}" if nums else None


def _shapes(module: torch.nn.Module) -> List[Tuple[str, torch.Size]]:
    return [(nm, p.shape) for nm, p in module.named_parameters()]


def save_checkpoint(trainer, directory, steps: int, epoch: int, final: bool = False) -> None:
    """Write the reference's three files for a GanTrainer (train.py:421-436 / :441-456)."""
    trainer.flush()
    torch.cuda.synchronize()
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    tag = "final" if final else f"{steps:08d}"
    cpu = lambda sd: OrderedDict((k, t.detach().cpu().clone()) for k, t in sd.items())
    torch.save(cpu(trainer.net_g.state_dict()), d / f"netG-{tag}.pt")
    torch.save(cpu(trainer.net_d.state_dict()), d / f"netD-{tag}.pt")
    # `initial_lr` is what torch.optim.lr_scheduler writes into every param group; the reference re-creates
    # ExponentialLR(..., last_epoch=start_epoch) right after loading (train.py:98-104), which REQUIRES the key
    hyper = dict(lr=trainer.lr, initial_lr=trainer.base_lr)
    opt = lambda net, fp: flat_to_adamw_state(_shapes(net), fp.offsets, fp.m, fp.v, int(fp.step.item()), hyper)
    torch.save(dict(epoch=epoch, steps=steps, optG=opt(trainer.net_g, trainer.G), optD=opt(trainer.net_d, trainer.D)),
               d / f"checkpoint-{tag}.pt")


def load_latest_checkpoint(trainer, directory) -> Tuple[int, int]:
    """common.py:23-61 for a GanTrainer: load the newest netG / netD / checkpoint triple (written by the reference loop or
    by save_checkpoint) into the flat parameter and moment buffers.  Returns (start_epoch, steps)."""
    d = Path(directory)
    latest = latest_step(d)
    if latest is None:
        raise FileNotFoundError(f"no checkpoint-*.pt in {d}")
    trainer.flush()
    dev = trainer.device
    # load_state_dict copies INTO the existing tensors, i.e. into the flat buffers the parameters are views of
    trainer.net_g.load_state_dict(fix_state_dict(torch.load(d / f"netG-{latest}.pt", map_location=dev)))
    trainer.net_d.load_state_dict(fix_state_dict(torch.load(d / f"netD-{latest}.pt", map_location=dev)))
    ck = torch.load(d / f"checkpoint-{latest}.pt", map_location="cpu")
    for net, fp, key in ((trainer.net_g, trainer.G, "optG"), (trainer.net_d, trainer.D, "optD")):
        m, v = torch.zeros(fp.numel), torch.zeros(fp.numel)
        step = adamw_state_to_flat(ck[key], _shapes(net), fp.offsets, m, v)
        fp.m.copy_(m); fp.v.copy_(v); fp.step.fill_(step)
        lr = ck[key]["param_groups"][0].get("lr")
        if lr is not None:
            trainer.lr = float(lr)                      # (ExponentialLR has been applied to it, train.py:98-104)
        if ck[key]["param_groups"][0].get("initial_lr") is not None:
            trainer.base_lr = float(ck[key]["param_groups"][0]["initial_lr"])
    # the discriminator's packed operands are stale, and a captured phase-D graph does not re-pack them (it reuses the
    # packs of the previous step's phase G): re-pack eagerly, now (G re-packs at the start of every step)
    trainer.refold()
    return int(ck["epoch"]), int(ck["steps"])
