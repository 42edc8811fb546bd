"""A communicator of our own on the NCCL library that PyTorch already loaded (ctypes, no torch types below the call):
`ncclAllReduce` is issued on an EXPLICIT CUDA stream, so the gradient exchange can be captured INSIDE the CUDA graphs of
the train step (NCCL supports stream capture) and ordered against the backward kernels with plain events - no host-side
seam between graph replays, no process-group work objects.  torch.distributed is still used for the rendezvous (the
128-byte ncclUniqueId travels through one broadcast), for barriers and for the initial parameter broadcast.

The reference is single-process (ste_gan/train.py:545); this is the exchange step that data parallelism adds.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
from typing import Optional

import torch

NCCL_FLOAT32, NCCL_SUM = 7, 0          # ncclDataType_t / ncclRedOp_t (nccl.h)


class NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


_lib = None


def load_library():
    """libnccl.so.2 - the copy bundled with torch (nvidia/nccl/lib) if present, else the system one."""
    global _lib
    if _lib is not None:
        return _lib
    cands = []
    spec = importlib.util.find_spec("nvidia.nccl") if importlib.util.find_spec("nvidia") else None
    if spec is not None and spec.submodule_search_locations:
        for loc in spec.submodule_search_locations:
            cands.append(os.path.join(loc, "lib", "libnccl.so.2"))
    cands += ["libnccl.so.2", "libnccl.so"]
    err = None
    for c in cands:
        try:
            lib = C.CDLL(c)
            break
        except OSError as e:      # noqa: PERF203
            err = e
    else:
        raise RuntimeError(f"libnccl not found: {err}")
    lib.ncclGetVersion.argtypes = [C.POINTER(C.c_int)]
    lib.ncclGetUniqueId.argtypes = [C.POINTER(NcclUniqueId)]
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, NcclUniqueId, C.c_int]
    lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    lib.ncclGetErrorString.argtypes = [C.c_int]
    lib.ncclGetErrorString.restype = C.c_char_p
    for f in ("ncclGetVersion", "ncclGetUniqueId", "ncclCommInitRank", "ncclAllReduce", "ncclCommDestroy"):
        getattr(lib, f).restype = C.c_int
    _lib = lib
    return lib


def version() -> int:
    v = C.c_int(0)
    _check(load_library().ncclGetVersion(C.byref(v)), "ncclGetVersion")
    return v.value


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what}: {load_library().ncclGetErrorString(rc).decode()} (ncclResult {rc})")


class NcclComm:
    """One communicator over the ranks of a torch.distributed group; the current CUDA device must already be set."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.lib = load_library()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        uid = NcclUniqueId()
        if self.rank == 0:
            _check(self.lib.ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
        on_cuda = dist.get_backend(group) == "nccl"
        t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device="cuda" if on_cuda else "cpu")
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        C.memmove(C.byref(uid), bytes(t.cpu().tolist()), 128)
        self.comm = C.c_void_p()
        _check(self.lib.ncclCommInitRank(C.byref(self.comm), self.world, uid, self.rank), "ncclCommInitRank")
        # first collective outside any capture: connection setup happens here
        warm = torch.zeros(8, device="cuda")
        self.all_reduce(warm)
        torch.cuda.synchronize()

    def all_reduce(self, t: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> None:
        """In-place SUM of a contiguous fp32 CUDA tensor over all ranks, enqueued on `stream` (default: current)."""
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise TypeError("NcclComm.all_reduce: contiguous fp32 CUDA tensor expected")
        s = (stream or torch.cuda.current_stream()).cuda_stream
        _check(self.lib.ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), NCCL_FLOAT32, NCCL_SUM, self.comm, s), "ncclAllReduce")

    def destroy(self) -> None:
        if self.comm:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = C.c_void_p()
