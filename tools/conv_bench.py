"""Micro-benchmark of single conv shapes through the C ABI (GPU): CUDA-graph replay of N back-to-back launches,
device-timed.   python tools/conv_bench.py [fwd|dgrad|wgrad|all] [reps] [shape-filter]
Shapes = the layer list of the bench configuration (B=16, 100 frames)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200 import ops

B = 16
# name, phases, t_in, c_in, c_out, k, dil, stride, pad, groups
SHAPES = [
    ("g768_T100_d3", 1, 100, 768, 768, 3, 3, 1, 3, 1),
    ("g768_T100_k1", 1, 100, 768, 768, 1, 1, 1, 0, 1),
    ("g384_T200_d9", 1, 200, 384, 384, 3, 9, 1, 9, 1),
    ("g384_T400_d1", 1, 400, 384, 384, 3, 1, 1, 1, 1),
    ("g384_T800_d27", 1, 800, 384, 384, 3, 27, 1, 27, 1),
    ("g192_T1600_d3", 1, 1600, 192, 192, 3, 3, 1, 3, 1),
    ("g384to192_T1600", 1, 1600, 384, 192, 3, 1, 1, 1, 1),
    ("s_k5_T400", 1, 400, 512, 1024, 5, 1, 1, 2, 1),
    ("s_k37g4_T1600", 1, 1600, 128, 256, 37, 1, 2, 18, 4),
    ("s_k37g16_T800", 1, 800, 256, 512, 37, 1, 2, 18, 16),
    ("f_k41g4_T1600", 1, 1600, 128, 128, 41, 1, 2, 20, 4),
    ("f_k41g16c8_T800", 1, 800, 128, 256, 41, 1, 2, 20, 16),
    ("f_k41g16s4_T400", 1, 400, 256, 512, 41, 1, 4, 20, 16),
    ("f_k41g16s4_T100", 1, 100, 512, 1024, 41, 1, 4, 20, 16),
    ("f_k41g16_T25", 1, 25, 1024, 1024, 41, 1, 1, 20, 16),
    ("s_out_T400", 1, 400, 1024, 1, 3, 1, 1, 1, 1),
    ("p11_l3", 11, 50, 256, 512, 3, 1, 3, 2, 1),
    ("p2_l3", 2, 269, 256, 512, 3, 1, 3, 2, 1),
    ("p2_l2", 2, 803, 32, 256, 3, 1, 3, 2, 1),
]


def t_out_of(T, k, d, s, pad):
    return (T + 2 * pad - d * (k - 1) - 1) // s + 1


def bench(fn, reps):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    dt = torch.bfloat16
    print(f"{'shape':18s} {'kind':6s} {'us':>8s} {'TFLOP/s':>8s}")
    for (name, p, T, ci, co, k, d, s, pad, g) in SHAPES:
        if filt and filt not in name:
            continue
        To = t_out_of(T, k, d, s, pad)
        pg = ops.tc_pack_groups(ci, co, g)
        x = torch.randn(B, T * p, ci, device="cuda").to(dt)
        dy = torch.randn(B, To * p, co, device="cuda").to(dt)
        wf = (torch.randn(k, co, ci // pg, device="cuda") / (ci // g * k) ** 0.5).to(dt)
        wd = (torch.randn(k, ci, co // pg, device="cuda") / (ci // g * k) ** 0.5).to(dt)
        bias = torch.randn(co, device="cuda")
        res = torch.randn(B, To * p, co, device="cuda").to(dt)
        y = torch.empty(B, To * p, co, device="cuda", dtype=torch.float32 if co < 8 else dt)
        ya = torch.empty(B, To * p, co, device="cuda", dtype=dt) if co >= 8 else None
        dx = torch.empty(B, T * p, ci, device="cuda", dtype=dt)
        ld, span = ops.wgrad_layout(dt, c_in=ci, c_out=co, k=k, groups=g, stride=s)
        dw = torch.zeros(co * ld, device="cuda")
        db = torch.zeros(co, device="cuda")
        flops = 2.0 * B * p * To * co * k * (ci // g)
        runs = []
        if which in ("fwd", "all"):
            runs.append(("fwd", lambda: ops.conv(x, wf, n_samples=B, phases=p, t_src=T, t_dst=To, c_src=ci, c_dst=co, groups=pg,
                                                  k=k, dilation=d, stride=s, pad=pad, bias=bias, act=ops.ACT_RELU,
                                                  add_post=res if co >= 8 else None, y_raw=y, y_act=ya)))
        if which in ("dgrad", "all") and co >= 8:
            # grouped convs: K-major data-gradient pack (compact groups); the others read the forward pack (MN-major)
            runs.append(("dgrad", lambda: ops.conv(dy, wd if g > 1 else wf, n_samples=B, phases=p, t_src=To, t_dst=T, c_src=co, c_dst=ci,
                                                    groups=pg, k=k, dilation=d, stride=s, pad=pad, transposed=True, mask=x,
                                                    mask_mode=ops.ACT_RELU, y_raw=dx, w_fwd_pack=g == 1)))
        if which in ("wgrad", "all") and co >= 32:
            runs.append(("wgrad", lambda: ops.wgrad(x, dy, dw, db, n_samples=B, phases=p, t_in=T, t_out=To, c_in=ci, c_out=co,
                                                    groups=g, k=k, dilation=d, stride=s, pad=pad)))
        for kind, fn in runs:
            us = bench(fn, reps)
            print(f"{name:18s} {kind:6s} {us:8.1f} {flops / us / 1e6:8.1f}", flush=True)


if __name__ == "__main__":
    main()
