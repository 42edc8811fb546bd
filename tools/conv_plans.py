"""Launch plans of the tcgen05 convolution engine for the layer classes of the benchmarked configurations (host-only, no GPU:
stg_debug_conv_plan).   python tools/conv_plans.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_host_cpu import _conv_plan
from ste_gan_b200 import ops

rows = []
def add(name, **kw):
    rc, pl = _conv_plan(**kw)
    rows.append((name, rc, pl))

for ci, co, k, dil, T in [(768, 768, 3, 3, 100), (768, 768, 1, 1, 100), (768, 384, 3, 1, 200), (384, 384, 3, 9, 400), (384, 384, 3, 27, 800),
                          (384, 192, 3, 1, 1600), (192, 192, 3, 3, 1600), (192, 192, 1, 1, 1600)]:
    geo = dict(n_samples=16, phases=1, t_src=T, t_dst=T, groups=1, k=k, dilation=dil, stride=1, pad=dil * (k - 1) // 2)
    add(f"G {ci}->{co} k{k} d{dil} T{T} fwd", **dict(geo, c_src=ci, c_dst=co, act=ops.ACT_RELU, ptrs=("y_raw", "y_act", "add_post", "bias")))
    add(f"G {ci}->{co} k{k} d{dil} T{T} dgrad", **dict(geo, c_src=co, c_dst=ci, transposed=1, w_fwd_pack=1, mask_mode=ops.ACT_RELU, ptrs=("y_raw", "mask")))
for ci, co, k, st, g, T, B in [(128, 256, 37, 2, 4, 1600, 16), (256, 512, 37, 2, 16, 800, 16), (512, 1024, 5, 1, 1, 400, 16), (512, 1024, 5, 1, 1, 200, 32),
                               (128, 128, 41, 2, 4, 1600, 16), (256, 512, 41, 4, 16, 400, 16), (1024, 1024, 41, 1, 16, 25, 16), (1024, 1024, 5, 1, 1, 25, 16)]:
    To = (T + 2 * (k // 2) - (k - 1) - 1) // st + 1
    geo = dict(n_samples=B, phases=1, k=k, dilation=1, stride=st, pad=k // 2, groups=ops.tc_pack_groups(ci, co, g))
    add(f"S {ci}->{co} k{k} s{st} g{g} T{T} B{B} fwd", **dict(geo, t_src=T, t_dst=To, c_src=ci, c_dst=co, act=ops.ACT_LEAKY, ptrs=("y_act", "bias")))
    add(f"S {ci}->{co} k{k} s{st} g{g} T{T} B{B} dgrad", **dict(geo, t_src=To, t_dst=T, c_src=co, c_dst=ci, transposed=1, w_fwd_pack=int(g == 1),
                                                               mask_mode=ops.ACT_LEAKY, ptrs=("y_raw", "mask")))
for p_, H, ci, co in [(2, 803, 32, 256), (2, 269, 256, 512), (11, 148, 32, 256), (11, 50, 256, 512)]:
    Ho = (H + 4 - 2 - 1) // 3 + 1
    geo = dict(n_samples=32, phases=p_, k=3, dilation=1, stride=3, pad=2, groups=1)
    add(f"P{p_} {ci}->{co} H{H} fwd", **dict(geo, t_src=H, t_dst=Ho, c_src=ci, c_dst=co, act=ops.ACT_LEAKY, ptrs=("y_act", "bias")))
    add(f"P{p_} {ci}->{co} H{H} dgrad", **dict(geo, t_src=Ho, t_dst=H, c_src=co, c_dst=ci, transposed=1, w_fwd_pack=1, mask_mode=ops.ACT_LEAKY, ptrs=("y_raw", "mask")))
cols = ("bn", "taps_per_stage", "stages", "smem", "occ2", "row_classes", "tiles", "grid", "compact_k", "k_chunks", "in_ring", "residues")
print(f"{'layer':44s} " + " ".join(f"{c:>8s}" for c in ("bn", "taps/stg", "stages", "smem KB", "2cta/sm", "row cls", "tiles", "grid", "cmpct K", "chunks", "in ring", "residues")))
for name, rc, pl in rows:
    if rc:
        print(f"{name:44s} unsupported ({rc})"); continue
    v = [pl[c] if c != "smem" else round(pl[c] / 1024) for c in cols]
    print(f"{name:44s} " + " ".join(f"{x:8d}" for x in v))
