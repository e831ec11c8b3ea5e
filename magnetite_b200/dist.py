"""Multi-GPU bootstrap: one process per GPU (torchrun), `torch.distributed` for the rendezvous.

The library owns its NCCL communicator (scalar allreduces inside the CG graph, one-off
exchanges) and maps the neighbours' halo buffers through CUDA IPC; torch.distributed only
carries the 128-byte NCCL unique id from rank 0 to the others and provides the barriers the
benchmark contract asks for.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: str = "nccl"):
    """torch.distributed rendezvous from the torchrun environment (MASTER_ADDR defaults to 127.0.0.1)."""
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    rank, world, local = rank_world()
    if not dist.is_initialized():
        if backend == "nccl":
            import torch
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_bytes(payload: bytes, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a small byte string from `src` with whatever backend the group uses."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def init_comm(ctx: _lib.Context) -> None:
    """Give the context an NCCL communicator spanning the torch.distributed world."""
    import torch.distributed as dist
    lib = _lib.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    uid = (C.c_ubyte * 128)()
    if rank == 0:
        _lib.check(lib.mag_comm_unique_id(C.cast(uid, C.c_void_p)), "mag_comm_unique_id")
    data = broadcast_bytes(bytes(uid), 128, src=0)
    uid2 = (C.c_ubyte * 128).from_buffer_copy(data)
    _lib.check(lib.mag_comm_init(ctx.handle, rank, world, C.cast(uid2, C.c_void_p)), "mag_comm_init")


def partition_nodes(n_nodes: int, nranks: int, rank: int):
    lo, hi = C.c_uint64(), C.c_uint64()
    _lib.check(_lib.load().mag_partition_nodes(n_nodes, nranks, rank, C.byref(lo), C.byref(hi)), "mag_partition_nodes")
    return int(lo.value), int(hi.value)


def halo_plan(nranks: int, rank: int, row_lo, ext_lo, ext_hi):
    """[(lo, hi, dst_rank)]: index ranges of `rank`'s rows that rank dst reads as halo (mag_halo_plan)."""
    row_lo = np.ascontiguousarray(row_lo, np.uint32); ext_lo = np.ascontiguousarray(ext_lo, np.uint32)
    ext_hi = np.ascontiguousarray(ext_hi, np.uint32)
    cap = 4 * nranks
    slo, shi, dst = np.empty(cap, np.uint32), np.empty(cap, np.uint32), np.empty(cap, np.int32)
    n = C.c_int32()
    _lib.check(_lib.load().mag_halo_plan(nranks, rank, _lib.ptr(row_lo), _lib.ptr(ext_lo), _lib.ptr(ext_hi),
                                         _lib.ptr(slo), _lib.ptr(shi), _lib.ptr(dst), cap, C.byref(n)), "mag_halo_plan")
    return [(int(slo[i]), int(shi[i]), int(dst[i])) for i in range(n.value)]
