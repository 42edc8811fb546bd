"""Do per-group tcgen05 MMAs work on narrow-swizzle weight tiles?  (GPU)  Prints rel-L2 per (cin_g, cout_g)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
for cin_g, cout_g in ((16, 32), (16, 16), (32, 64), (32, 32), (8 * 2, 64), (64, 64), (64, 128), (32, 128)):
    n_g = 64 // cin_g
    x = torch.randn(128, 64).bfloat16().cuda(); w = torch.randn(n_g * cout_g, cin_g).bfloat16().cuda()
    ref = torch.cat([x[:, q * cin_g:(q + 1) * cin_g].float() @ w[q * cout_g:(q + 1) * cout_g].float().t() for q in range(n_g)], 1)
    out = torch.full((128, n_g * cout_g), float("nan"), device="cuda")
    rc = lib.stg_debug_group_mma(x.data_ptr(), w.data_ptr(), cin_g, cout_g, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print(f"cin_g {cin_g:3d} cout_g {cout_g:3d}: rc={rc} rel-L2 {float((out - ref).norm() / ref.norm()):.3e}")
