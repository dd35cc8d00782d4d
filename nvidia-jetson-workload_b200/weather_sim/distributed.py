"""Row-slab domain decomposition helpers: one process per GPU, rendezvous through torch.distributed.

torch.distributed is plumbing only (rank discovery, broadcasting the 128-byte NCCL id, barriers and the
max-over-ranks reduction of timings). The data path -- ghost-row exchange between neighbouring slabs -- is
ncclSend/ncclRecv issued by libweather_b200.so itself on its own comm stream (csrc/wsb_nccl.cpp).
"""
import os

import numpy as np

from . import _capi


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def slab_rows(grid_height, nranks, rank):
    """(row0, nrows) owned by `rank`; identical to the C library's partition (wsb_partition_rows)."""
    return _capi.partition_rows(grid_height, nranks, rank)


def local_slab(global_array, nranks, rank):
    r0, n = slab_rows(global_array.shape[-2], nranks, rank)
    return np.ascontiguousarray(global_array[..., r0:r0 + n, :])


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from `src` to every rank of the default process group."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def share_nccl_id():
    """Rank 0 creates the NCCL unique id through the C-ABI; every rank returns the same 128 bytes."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    ident = _capi.nccl_unique_id() if dist.get_rank() == 0 else None
    return broadcast_bytes(ident, 0)


def slab_simulation(width, grid_height, rank, nranks, device_id, nccl_id=None, **kw):
    """Simulation handle for this rank's slab of a (grid_height x width) global grid."""
    if nranks > 1 and nccl_id is None:
        nccl_id = share_nccl_id()
    return _capi.Simulation(width, grid_height, rank=rank, nranks=nranks, nccl_id=nccl_id, device_id=device_id, **kw)


def gather_rows(local, dst=0):
    """Concatenate every rank's slab on `dst` (tests / small grids only)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, parts, dst=dst)
    if dist.get_rank() == dst:
        return np.concatenate(parts, axis=-2)
    return None
