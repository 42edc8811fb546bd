"""All-reduce cost of the step's gradient pieces through the own NCCL communicator, alone on the chip (N ranks):
    torchrun --nproc-per-node N tools/nccl_probe.py      (env: NCCL_ALGO / NCCL_PROTO / NCCL_MAX_NCHANNELS ... to compare)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ste_gan_b200.dist import init_from_env
from ste_gan_b200.nccl import NcclComm

rank, world, local = init_from_env("nccl")
torch.cuda.set_device(local)
comm = NcclComm()
sizes_mb = [47.4, 31.0, 31.0, 31.0, 11.7, 4.0, 1.0]
bufs = [torch.ones(int(mb * 1e6 / 4), device="cuda") for mb in sizes_mb]
for b in bufs:
    comm.all_reduce(b)
torch.cuda.synchronize(); dist.barrier()
out = []
for mb, b in zip(sizes_mb, bufs):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for _ in range(10):
            comm.all_reduce(b)
    g.replay(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 10
    out.append(f"{mb:5.1f} MB {us:7.1f} us {mb * 1e6 / us / 1e3:6.1f} GB/s")
if rank == 0:
    print(f"world {world} " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NCCL_")) + "\n  " + "\n  ".join(out), flush=True)
dist.barrier()
dist.destroy_process_group()
