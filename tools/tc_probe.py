"""Diagnostic probe for the tcgen05 engines (run on the GPU box; prints to stdout).

Feeds impulses through the tensor-core conv / wgrad and reports where they land, plus
error statistics for a few dense cases, so that one GPU trip can discriminate between
descriptor / swizzle / addressing hypotheses.
"""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200 import ops
from tests.refconv import conv_ref, pack_fwd, pack_dgrad, rel_l2


def fwd(x, w, k, d, pad, engine, stride=1, phases=1):
    B, TP, ci = x.shape
    T = TP // phases
    co = w.shape[0]
    To = (T + 2 * pad - d * (k - 1) - 1) // stride + 1
    y = torch.full((B, To * phases, co), float("nan"), device="cuda", dtype=torch.float32)
    ops.conv(x.cuda().bfloat16(), pack_fwd(w).cuda().bfloat16(), n_samples=B, phases=phases, t_src=T, t_dst=To,
             c_src=ci, c_dst=co, k=k, dilation=d, stride=stride, pad=pad, y_raw=y, engine=engine)
    torch.cuda.synchronize()
    return y.cpu()


def impulse():
    print("== impulse probe: x[0,t0,c0]=1, w[co,ci,0]=co*1000+ci (k=1) -> y[0,t0,:] should be co*1000+c0")
    for (T, ci, co, t0, c0) in [(128, 64, 128, 5, 3), (128, 64, 128, 77, 40), (128, 128, 128, 9, 100), (200, 64, 16, 150, 17)]:
        x = torch.zeros(1, T, ci); x[0, t0, c0] = 1.0
        w = (torch.arange(co).view(co, 1, 1) * 1.0 + torch.arange(ci).view(1, ci, 1) / 256.0)
        y = fwd(x, w, 1, 1, 0, ops.ENGINE_TCGEN05)
        nz = (y.abs() > 0).nonzero()
        rows = sorted(set(nz[:, 1].tolist()))
        exp = w[:, c0, 0].bfloat16().float()
        ok = torch.allclose(y[0, t0], exp)
        print(f"T={T} ci={ci} co={co} t0={t0} c0={c0}: nonzero rows={rows[:8]} n_nz={len(nz)} nan={int(y.isnan().sum())} row_ok={ok}")
        if not ok:
            print("   got ", y[0, t0, :8].tolist()); print("   want", exp[:8].tolist())
            if len(rows):
                r = rows[0]; print("   row", r, y[0, r, :8].tolist())


def dense():
    print("== dense cases (rel L2 vs fp64 reference on bf16-rounded inputs)")
    gen = torch.Generator().manual_seed(0)
    for (B, T, ci, co, k, d, s, p) in [(1, 128, 64, 128, 1, 1, 1, 1), (2, 100, 64, 128, 3, 1, 1, 1), (2, 100, 128, 64, 3, 27, 1, 1),
                                    (1, 100, 768, 384, 3, 3, 1, 1), (2, 300, 192, 192, 3, 9, 1, 1), (2, 160, 192, 8, 3, 1, 1, 1),
                                    (2, 200, 64, 128, 3, 1, 2, 1), (2, 140, 256, 128, 3, 1, 3, 2), (1, 50, 512, 128, 5, 1, 1, 1)]:
        try:
            x = torch.randn(B, T * p, ci, generator=gen).bfloat16().float()
            w = (torch.randn(co, ci, k, generator=gen) / (ci * k) ** 0.5).bfloat16().float()
            pad = d * (k - 1) // 2
            y = fwd(x, w, k, d, pad, ops.ENGINE_TCGEN05, stride=s, phases=p)
            ref = conv_ref(x, w, None, phases=p, stride=s, dilation=d, pad=pad)
            ys = fwd(x, w, k, d, pad, ops.ENGINE_SIMT, stride=s, phases=p)
            print(f"B={B} T={T} ci={ci} co={co} k={k} d={d} s={s} p={p}: tc {rel_l2(y, ref):.3e}  simt {rel_l2(ys, ref):.3e}  nan={int(y.isnan().sum())}")
        except Exception as e:
            print("case failed:", (B, T, ci, co, k, d, s, p), repr(e))


def wgrad_probe():
    print("== wgrad impulse: dy[0,t0,co0]=1, x[0,t0,:]=arange -> dw[co0,0,:] should be arange (k=1)")
    for (T, ci, co, t0, co0) in [(64, 64, 128, 5, 3), (128, 128, 128, 70, 100), (100, 192, 192, 33, 150)]:
        try:
            x = torch.zeros(1, T, ci); x[0, t0] = torch.arange(ci).float()
            dy = torch.zeros(1, T, co); dy[0, t0, co0] = 1.0
            dw = torch.zeros(co, 1, ci, device="cuda")
            ops.wgrad(x.cuda().bfloat16(), dy.cuda().bfloat16(), dw, None, n_samples=1, t_in=T, t_out=T, c_in=ci, c_out=co,
                      k=1, engine=ops.ENGINE_TCGEN05)
            torch.cuda.synchronize()
            dw = dw.cpu()
            nzr = sorted(set((dw.abs() > 0).nonzero()[:, 0].tolist()))
            ok = torch.allclose(dw[co0, 0], torch.arange(ci).float())
            print(f"T={T} ci={ci} co={co} t0={t0} co0={co0}: nonzero co rows={nzr[:8]} ok={ok}")
            if not ok and nzr:
                print("   row", nzr[0], dw[nzr[0], 0, :12].tolist())
        except Exception as e:
            print("wgrad case failed:", repr(e))
    gen = torch.Generator().manual_seed(1)
    for (B, T, ci, co, k, d, s) in [(2, 100, 64, 128, 3, 1, 1), (2, 300, 192, 192, 3, 9, 1), (1, 100, 768, 384, 3, 3, 1), (2, 200, 64, 128, 3, 1, 2)]:
        try:
            pad = d * (k - 1) // 2
            To = (T + 2 * pad - d * (k - 1) - 1) // s + 1
            x = torch.randn(B, T, ci, generator=gen).bfloat16().float()
            dy = torch.randn(B, To, co, generator=gen).bfloat16().float()
            outs = []
            for eng in (ops.ENGINE_TCGEN05, ops.ENGINE_SIMT):
                dw = torch.zeros(co, k, ci, device="cuda")
                ops.wgrad(x.cuda().bfloat16(), dy.cuda().bfloat16(), dw, None, n_samples=B, t_in=T, t_out=To, c_in=ci,
                          c_out=co, k=k, dilation=d, stride=s, pad=pad, engine=eng)
                torch.cuda.synchronize(); outs.append(dw.cpu())
            wr = torch.zeros(co, ci, k, dtype=torch.double, requires_grad=True)
            yy = conv_ref(x, wr, None, stride=s, dilation=d, pad=pad)
            (gw,) = torch.autograd.grad(yy, wr, dy.double())
            gw = gw.permute(0, 2, 1)
            print(f"wgrad B={B} T={T} ci={ci} co={co} k={k} d={d} s={s}: tc {rel_l2(outs[0], gw):.3e} simt {rel_l2(outs[1], gw):.3e}")
        except Exception as e:
            print("wgrad dense failed:", repr(e))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for fn in (impulse, dense, wgrad_probe):
        try:
            fn()
        except Exception:
            traceback.print_exc()
