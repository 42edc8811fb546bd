"""What one SM ingests through TMA while N SMs pull L2-resident boxes at once (GPU).  python tools/tma_bw_probe.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200 import _lib
lib = _lib.load()
n_rows = 48 * 1024 * 1024 // 128          # 48 MB: stays in L2
buf = torch.randn(n_rows, 64, device="cuda").to(torch.bfloat16)
print(f"{'grid':>5s} {'box_rows':>8s} {'stages':>6s} {'mode':>9s} {'B/clk/SM':>9s} {'min':>7s} {'max':>7s} {'TB/s chip':>9s}")
for grid in (1, 37, 74, 148):
    for box_rows, stages in ((128, 8), (256, 6), (64, 8)):
        for mode in (0, 1):
            n_iters = 48 * 1024 * 1024 // (box_rows * 128) // 148       # distinct mode: one pass over the buffer at 148 CTAs
            clk = torch.zeros(grid, dtype=torch.int64, device="cuda")
            for _ in range(3):
                rc = lib.stg_debug_tma_bw(buf.data_ptr(), n_rows, box_rows, n_iters, mode, stages, grid, clk.data_ptr(),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream))
                assert rc == 0, rc
                torch.cuda.synchronize()
            c = clk.double()
            per = n_iters * box_rows * 128 / c
            print(f"{grid:5d} {box_rows:8d} {stages:6d} {'same' if mode else 'distinct':>9s} {per.mean():9.1f} {per.min():7.1f} {per.max():7.1f} "
                  f"{per.sum().item() * 1.965e9 / 1e12:9.2f}")
