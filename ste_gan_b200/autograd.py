"""torch.autograd.Function wrappers around the fused passes, so that the drop-in modules work
with the reference's unmodified train loop (`loss.backward()`, `optimizer.step()`,
ste_gan/train.py:165-268).  The fast path (ste_gan_b200/trainer.py) calls the passes
directly and never goes through autograd.

The passes accumulate parameter gradients into `.grad` themselves; inside an autograd
backward they must be *returned* instead, so `_collect_param_grads` runs the pass against
temporarily detached `.grad` slots and hands the results to the autograd engine.
"""
from __future__ import annotations

from typing import List

import torch

from . import ops, passes
from .precision import act_dtype


class _collect_param_grads:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params

    def __enter__(self):
        self.saved = [p.grad for p in self.params]
        for p in self.params:
            p.grad = None
        return self

    def __exit__(self, *exc):
        self.grads = [p.grad for p in self.params]
        for p, g in zip(self.params, self.saved):
            p.grad = g
        return False


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: ste_gan_b200 runs on CUDA (sm_100a) only - there is no CPU path; got a {t.device} tensor")


class GeneratorFn(torch.autograd.Function):
    """x_pred = G(units, session_ids, speaking_mode_ids) - models/generator.py:140-162."""

    @staticmethod
    def forward(ctx, model, units, session_ids, speaking_mode_ids, *params):
        _require_cuda(units, "EMGGeneratorGanTTS.forward")
        need = any(ctx.needs_input_grad[4:])
        x_pred, gctx = passes.generator_forward(model, units, session_ids, speaking_mode_ids, act_dtype(), need)
        ctx.model, ctx.gctx, ctx.n_params = model, gctx, len(params)
        return x_pred

    @staticmethod
    def backward(ctx, grad_out):
        params = list(ctx.model.parameters())
        with _collect_param_grads(params) as col:
            passes.generator_backward(ctx.model, ctx.gctx, grad_out.contiguous().float())
        ctx.gctx = None
        return (None, None, None, None) + tuple(col.grads)


class DiscriminatorFn(torch.autograd.Function):
    """List[List[fmap]] = D(x) - models/discriminator.py:144-155,180-191 (flattened to a tuple)."""

    @staticmethod
    def forward(ctx, model, x, *params):
        _require_cuda(x, "Discriminator.forward")
        dtype = act_dtype()
        folds = passes.fold_discriminator(model, dtype, training=model.training)
        results, dctx = passes.discriminator_forward(model, x, dtype, folds)
        ctx.model, ctx.dctx = model, dctx
        ctx.layout = [(s["kind"], s["phases"], len(r)) for s, r in zip(dctx.subs, results)]
        outs = []
        for sub, fmaps in zip(dctx.subs, results):
            outs += [passes.to_reference_layout(fm, sub["kind"], sub["phases"]) for fm in fmaps]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        dctx = ctx.dctx
        dlogits, dfmaps, i = [], [], 0
        for kind, phases, n in ctx.layout:
            gs = []
            for j in range(n):
                g = grads[i + j]
                if g is not None:
                    g = (g.permute(0, 2, 3, 1) if kind == "P" else g.transpose(1, 2)).contiguous()
                    g = g.reshape(g.shape[0], -1, g.shape[-1]).to(dctx.dtype)
                gs.append(g)
            i += n
            dlogits.append(gs[-1]); dfmaps.append(gs[:-1])
        params = list(ctx.model.parameters())
        want_w = any(ctx.needs_input_grad[2:])
        with _collect_param_grads(params) as col:
            dx = passes.discriminator_backward(ctx.model, dctx, dlogits, dfmaps, want_input_grad=ctx.needs_input_grad[1],
                                               want_weight_grad=want_w)
        ctx.dctx = None
        return (None, dx) + tuple(col.grads)


_TD_RES = ((20, 8), (51, 13), (80, 16))        # time_domain_loss.py:88-93


class TdLossFn(torch.autograd.Function):
    """losses[n_res] = TD(x_real, x_gen) for a list of (win, shift) resolutions - losses/time_domain_loss.py:70-73,96-103.
    The backward hands the upstream gradients to the kernels as a DEVICE array: no host synchronisation."""

    @staticmethod
    def forward(ctx, x_real, x_gen, resolutions, pad_windows, avg_window):
        _require_cuda(x_gen, "TimeDomainFeatureLoss")
        xr, xg = x_real.detach().contiguous().float(), x_gen.detach().contiguous().float()
        losses = torch.zeros(len(resolutions), device=xg.device, dtype=torch.float32)
        ops.td_loss_ex(xr, xg, resolutions, losses, pad_windows=pad_windows, avg_window=avg_window)
        ctx.save_for_backward(xr, xg)
        ctx.cfg = (tuple(resolutions), pad_windows, avg_window)
        return losses

    @staticmethod
    def backward(ctx, g):
        xr, xg = ctx.saved_tensors
        res, pad_windows, avg_window = ctx.cfg
        # the partial losses share one backward; their upstream gradients are applied per resolution, on the device
        dx = torch.zeros_like(xg)
        scratch = torch.zeros(len(res), device=xg.device, dtype=torch.float32)
        ops.td_loss_ex(xr, xg, res, scratch, pad_windows=pad_windows, avg_window=avg_window,
                       grad_scale=g.detach().contiguous().float(), dx_gen=dx)
        return None, dx, None, None, None


class MultiTdLossFn:
    """(loss_20_8, loss_51_13, loss_80_16) = TD(x_real, x_gen) - losses/time_domain_loss.py:96-103."""

    @staticmethod
    def apply(x_real, x_gen):
        return TdLossFn.apply(x_real, x_gen, _TD_RES, True, 9)


class EncoderFn(torch.autograd.Function):
    """(speech units, phoneme logits) = frozen EMG encoder(x) - models/emg_encoder.py:71-88 in eval mode; differentiable
    w.r.t. the EMG input only (the encoder's weights are frozen, losses/emg_encoder_loss.py:61)."""

    @staticmethod
    def forward(ctx, enc, x):
        from . import passes_encoder as pe
        _require_cuda(x, "EMGEncoderTransformer.forward")
        plan = enc.plan(act_dtype())
        units, logits, ectx = pe.encoder_forward(plan, x.detach(), need_ctx=ctx.needs_input_grad[1])
        ctx.plan, ctx.ectx, ctx.in_dtype = plan, ectx, x.dtype
        return units, logits

    @staticmethod
    def backward(ctx, g_units, g_logits):
        from . import passes_encoder as pe
        dx = pe.encoder_backward(ctx.plan, ctx.ectx, g_units, g_logits)
        ctx.ectx = None
        return None, dx.to(ctx.in_dtype)


class EncoderLossFn(torch.autograd.Function):
    """(speech-unit loss, phoneme loss) of losses/emg_encoder_loss.py:63-84 from the encoder's two heads, value and gradient
    in one fused kernel (the upstream gradients reach it as device scalars: no host synchronisation)."""

    @staticmethod
    def forward(ctx, unit_pred, unit_target, logits, phoneme_target):
        _require_cuda(unit_pred, "EMGEncoderLoss")
        up, lg = unit_pred.detach().contiguous().float(), logits.detach().contiguous().float()
        slots = torch.zeros(2, device=up.device, dtype=torch.float32)
        # gradients for unit upstream scales (they are linear in the upstream gradient: scaled in backward)
        du, dl = ops.encoder_losses(up, unit_target.detach(), lg, phoneme_target, slots, 1.0, 1.0, torch.float32)
        ctx.save_for_backward(du, dl)
        return slots[0], slots[1]

    @staticmethod
    def backward(ctx, g_su, g_ph):
        du, dl = ctx.saved_tensors
        return du * g_su, None, dl * g_ph, None


class SingleConvFn(torch.autograd.Function):
    """One normalised conv layer in the reference layout ([B,C,T] or [B,C,H,W]) - layers/conv.py:16-17,89-101."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        _require_cuda(x, type(mod).__name__)
        dtype = act_dtype()
        two_d = x.dim() == 4
        if two_d:
            B, C, H, W = x.shape
            xcl = x.permute(0, 2, 3, 1).reshape(B, H * W, C).contiguous().to(dtype)
            phases, t = W, H
        else:
            B, C, T = x.shape
            xcl = x.transpose(1, 2).contiguous().to(dtype)
            phases, t = 1, T
        f = passes.fold(mod, dtype, training=mod.training)
        if f.unfold:
            xcl = passes.unfold_input(f, xcl, B, t, phases)
        y, _, t_out = passes._fwd(f, xcl, B, t, phases=phases, want_raw=True)
        ctx.mod, ctx.f, ctx.xcl, ctx.geom, ctx.in_dtype = mod, f, xcl, (B, t, t_out, phases, two_d), x.dtype
        y = y.to(x.dtype if x.dtype != torch.float64 else torch.float32)
        return passes.to_reference_layout(y, "P" if two_d else "S", phases)

    @staticmethod
    def backward(ctx, gy):
        B, t, t_out, phases, two_d = ctx.geom
        mod, f = ctx.mod, ctx.f
        g = (gy.permute(0, 2, 3, 1) if two_d else gy.transpose(1, 2)).contiguous()
        g = g.reshape(B, -1, g.shape[-1]).to(f.dtype)
        params = list(mod.parameters())
        with _collect_param_grads(params) as col:
            ws = passes._Workspace([f], g.device)
            passes._wgrad(f, ctx.xcl, g, B, t, t_out, ws, phases=phases)
        dx = None
        if ctx.needs_input_grad[1]:
            dx = passes._dgrad(f, g, B, t_out, t, phases=phases)
            dx = passes.to_reference_layout(dx, "P" if two_d else "S", phases).to(ctx.in_dtype)
        return (None, dx) + tuple(col.grads)


class GBlockFn(torch.autograd.Function):
    """One GBlock in the reference layout [B,C,T] - layers/conv.py:82-84."""

    @staticmethod
    def forward(ctx, blk, x, *params):
        _require_cuda(x, "GBlock")
        dtype = act_dtype()
        B, C, T = x.shape
        x_raw = x.transpose(1, 2).contiguous().to(dtype)
        # stand-alone layer API only; inside the models relu / upsample are produced by the previous convolution's epilogue
        if blk.upsample not in (1, 2):
            raise ValueError("GBlock upsample must be 1 or 2 on this path")
        x_act = ops.relu_rows(x_raw, blk.upsample > 1)
        folds = {id(c): passes.fold(c, dtype) for c in blk.convs().values()}
        y, _, s = passes.gblock_fwd(blk, folds, x_raw, x_act, B, T, want_raw=True, want_act=False,
                                    dup_next=False)
        ctx.blk, ctx.folds, ctx.saved, ctx.B, ctx.in_dtype = blk, folds, s, B, x.dtype
        return y.to(x.dtype).transpose(1, 2)

    @staticmethod
    def backward(ctx, gy):
        blk = ctx.blk
        dy = gy.transpose(1, 2).contiguous().to(next(iter(ctx.folds.values())).dtype)
        params = list(blk.parameters())
        with _collect_param_grads(params) as col:
            ws = passes._Workspace(list(ctx.folds.values()), dy.device)
            dx = passes.gblock_bwd(blk, ctx.folds, ctx.saved, dy, ctx.B, ws)
        return (None, dx.to(ctx.in_dtype).transpose(1, 2)) + tuple(col.grads)
