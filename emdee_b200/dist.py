"""Multi-GPU plumbing (one process per GPU): rendezvous through torch.distributed, data path through the
library's own NCCL communicator (EmDeeX_comm_init). PyTorch is only the launcher-side plumbing here: it
carries the 128-byte NCCL id from rank 0 to the other ranks, barriers, and reduces timings.

Works with the `gloo` backend on CPU for everything except the device communicator itself, which is what the
world_size-2 CPU tests exercise (slab arithmetic, id broadcast, result gathering)."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Tuple

import numpy as np


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def slab_range(M: int, rank: int, world: int) -> Tuple[int, int]:
    """Cell layers [z0, z1) of `rank` -- must equal EmDeeX_slab_range (engine.cu: slab_range)."""
    return (rank * M) // world, ((rank + 1) * M) // world


def cells_per_dim(L: float, rc: float, skin: float) -> int:
    """reference neighbor_lists.f90:185-186: M = max(floor(ndiv*L/xRc), 2*ndiv+1), ndiv = 2."""
    return max(int(np.floor(2.0 * L / (rc + skin))), 5)


def broadcast_bytes(payload: bytes, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a fixed-size byte string from `src` with whatever backend the process group uses."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.tensor(list(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def init_comm(lib, system) -> None:
    """Give `system` (already created on every rank, nothing uploaded yet) its NCCL communicator."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    buf = C.create_string_buffer(128)
    if rank == 0:
        lib.EmDeeX_comm_unique_id(buf)
    uid = broadcast_bytes(buf.raw, 128, 0)
    lib.EmDeeX_comm_init(system.md, rank, world, uid)


def gather_pairs(local_pairs: np.ndarray) -> np.ndarray:
    """All ranks' neighbor pairs (each rank exports the pairs whose lower-index atom it owns) -> sorted set."""
    import torch.distributed as dist
    world = dist.get_world_size()
    if world == 1:
        return local_pairs
    parts: List[np.ndarray] = [None] * world
    dist.all_gather_object(parts, local_pairs)
    allp = np.concatenate([p.reshape(-1, 2) for p in parts], axis=0)
    order = np.lexsort((allp[:, 1], allp[:, 0]))
    return allp[order]


def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
