"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only carries the 128-byte NCCL id.

The data path has ONE exchange step, inside the C library (csrc/annembed_cuda.cu): an in-place ncclAllGather
of the rows each rank owns every few sweeps, under the next launch; with peer memory (exchange_layout_handles) the moves
of nodes owned elsewhere are reduced straight into the owner's replica over NVLink.
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(n: int, rank: int, nranks: int) -> tuple[int, int]:
    """Owned node range of `rank` in the library's INTERNAL node order (set_shard() in csrc/annembed_cuda.cu; the internal
    order is a relabelling, so this is not a range of the caller's ids)."""
    n_pad = ((n + nranks - 1) // nranks + 31) // 32 * 32          # whole warp tiles per shard
    return min(n, rank * n_pad), min(n, (rank + 1) * n_pad)


def env_rank_world() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def broadcast_unique_id(make_id, rank: int, nranks: int) -> np.ndarray | None:
    """Rank 0 calls make_id() (-> 128 uint8) and every rank receives it through torch.distributed
    (any backend: gloo on CPU, nccl on GPUs).  Returns None when nranks == 1."""
    if nranks == 1:
        return None
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("torch.distributed must be initialised before broadcast_unique_id")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = np.ascontiguousarray(make_id(), np.uint8)
        assert uid.shape == (128,)
        t.copy_(torch.from_numpy(uid))
    dist.broadcast(t, src=0)
    return t.cpu().numpy().copy()


def exchange_layout_handles(ctx, rank: int, nranks: int) -> None:
    """Peer-memory set-up: all-gather the CUDA IPC handles of every rank's layout buffers (2 x 64 bytes each)
    through torch.distributed and open them in `ctx` (annembed_cuda_comm_import_layouts)."""
    if nranks == 1:
        return
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(ctx.export_layout()).to(dev)
    allh = torch.zeros(nranks * 128, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh, mine)
    ctx.import_layouts(allh.cpu().numpy())
