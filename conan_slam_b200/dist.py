"""Host-side rendezvous helpers for the multi-GPU paths (one process per GPU).

The library itself only needs a 128-byte NCCL unique id on every rank; how it travels is the
host's business.  Here it is broadcast with torch.distributed (any backend: nccl on the GPU box,
gloo in the CPU tests), which is also what provides barriers and max-over-ranks timing in bench.py.
"""
import ctypes as C

from . import _lib


def nccl_unique_id_local():
    """128-byte id created by NCCL on this process (call on rank 0)."""
    lib = _lib.load_library()
    buf = C.create_string_buffer(128)
    _lib.check(lib.cslam_nccl_unique_id(C.cast(buf, C.c_void_p)), "cslam_nccl_unique_id")
    return buf.raw


def broadcast_bytes(payload, nbytes, src=0, device=None):
    """Broadcast a fixed-size byte string from `src` over the default torch.distributed group."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def nccl_unique_id(device=None):
    """Same 128-byte NCCL id on every rank of the default process group."""
    import torch.distributed as dist
    payload = nccl_unique_id_local() if dist.get_rank() == 0 else b""
    return broadcast_bytes(payload, 128, src=0, device=device)


def shard_owner(row, world):
    """Rank that stores covariance row `row` (block-cyclic 128-row tiles, csrc/common.cuh Shard)."""
    return (row >> 7) % world


def shard_local_row(row, world):
    return ((row >> 7) // world) * 128 + (row & 127)


def shard_rows(n, rank, world):
    """Global row indices stored on `rank` for an n x n covariance."""
    return [i for i in range(n) if shard_owner(i, world) == rank]
