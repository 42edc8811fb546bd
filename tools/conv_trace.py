"""In-kernel timeline of CTA 0 of one tcgen05 conv launch (GPU).  python tools/conv_trace.py <shape-name> [fwd|dgrad]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200 import ops, _lib
from tools.conv_bench import SHAPES, t_out_of, B

name = sys.argv[1]; kind = sys.argv[2] if len(sys.argv) > 2 else "fwd"
(_, p, T, ci, co, k, d, s, pad, g) = next(x for x in SHAPES if x[0] == name)
dt = torch.bfloat16
To = t_out_of(T, k, d, s, pad); pg = ops.tc_pack_groups(ci, co, g)
x = torch.randn(B, T * p, ci, device="cuda").to(dt); dy = torch.randn(B, To * p, co, device="cuda").to(dt)
wf = (torch.randn(k, co, ci // pg, device="cuda") / (ci // g * k) ** 0.5).to(dt)
wd = (torch.randn(k, ci, co // pg, device="cuda") / (ci // g * k) ** 0.5).to(dt)
bias = torch.randn(co, device="cuda"); res = torch.randn(B, To * p, co, device="cuda").to(dt)
y = torch.empty(B, To * p, co, device="cuda", dtype=dt); ya = torch.empty_like(y); dx = torch.empty(B, T * p, ci, device="cuda", dtype=dt)
def run():
    if kind == "fwd":
        ops.conv(x, wf, n_samples=B, phases=p, t_src=T, t_dst=To, c_src=ci, c_dst=co, groups=pg, k=k, dilation=d, stride=s,
                 pad=pad, bias=bias, act=ops.ACT_RELU, add_post=res, y_raw=y, y_act=ya)
    else:
        ops.conv(dy, wf, n_samples=B, phases=p, t_src=To, t_dst=T, c_src=co, c_dst=ci, groups=pg, k=k, dilation=d, stride=s,
                 pad=pad, transposed=True, mask=x, mask_mode=ops.ACT_RELU, y_raw=dx, w_fwd_pack=True)
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = torch.zeros(1 + 3 * 5300, device="cuda", dtype=torch.int64)
lib = _lib.load()
lib.stg_debug_set_trace(buf.data_ptr())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(4):
        run()
buf.zero_(); torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
lib.stg_debug_set_trace(None)
life = [e for e in buf[1 + 3 * 3900:1 + 3 * 3900 + 3 * 16].view(-1, 3).cpu().tolist() if e[0] != 0]
lt0 = life[0][2]
ln = {30: "entry", 31: "set-up done", 32: "dependency released", 33: "exit"}
print("life cycle of CTA 0 over 4 back-to-back launches (one CUDA graph, PDL):")
for tag, val, t in life:
    print(f"{(t - lt0) / 1e3:9.2f} us  {ln.get(tag, tag)}")
print("last launch, CTA 0:")
buf[1 + 3 * 3900:1 + 3 * 3900 + 3 * 16] = 0
ev = [e for e in buf[1:1 + 3 * 3990].view(-1, 3).cpu().tolist() if e[0] != 0]
ev.sort(key=lambda e: e[2]); t0 = ev[0][2]
names = {4: "  stage free (producer)", 5: "  stage landed (mma)", 1: "producer tile", 2: "mma start", 3: "mma issued", 10: "epi tile start", 11: "epi acc ready", 12: "epi sub done", 20: "  sub: acc in regs", 21: "  sub: inputs landed", 22: "  sub: smem written", 23: "  sub: out slot free", 24: "  sub: barrier passed"}
for tag, val, t in ev[:int(os.environ.get("TRACE_N", "120"))]:
    print(f"{(t - t0) / 1e3:9.2f} us  {names.get(tag, tag):16s} {val}")

pr = buf[1 + 3 * 5200:1 + 3 * 5200 + 12].cpu().tolist()
if pr[3] or pr[10]:
    n = max(pr[3], 1); m = max(pr[10], 1)
    print(f"PROF producer: {n} stages, clk/stage: empty-wait {pr[0]/n:.0f}  expect+A {pr[1]/n:.0f}  W+advance {pr[2]/n:.0f}")
    print(f"PROF mma     : {m} stages, clk/stage: full-wait {pr[8]/m:.0f}  issue+commit {pr[9]/m:.0f}")
