"""Does tcgen05 accept row-shifted K-major SW128 operand views?  (GPU)  Prints rel-L2 per (shift, base_offset mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ste_gan_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
rows_a = 192
x = torch.randn(rows_a, 64).bfloat16().cuda(); w = torch.randn(64, 64).bfloat16().cuda()
for shift in (0, 1, 2, 3, 5, 7, 8, 9, 27, 54):
    ref = x[shift:shift + 128].float() @ w.float().t()
    res = []
    for mode, bo in (("base_off=0", 0), ("base_off=shift&7", shift & 7)):
        out = torch.full((128, 64), float("nan"), device="cuda")
        rc = lib.stg_debug_rowshift(x.data_ptr(), w.data_ptr(), rows_a, shift, bo, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        err = float((out - ref).norm() / ref.norm())
        res.append(f"{mode}: rc={rc} err={err:.3e}")
    print(f"shift {shift:3d}: " + " | ".join(res))
