"""Data parallelism over text lines: one process per GPU, lines sharded across ranks,
ONE start-up broadcast of the glyph index (and candidate chars) from rank 0, a gather of {key: text} at the end.
There is no steady-state collective: lines are independent (infer_effocr.py:549-565 loops over them one by one).

Works with backend "nccl" (one rank per B200, tensors on the device) and "gloo" (CPU tensors; used by the
CPU tests of this module).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun-style initialisation; returns (rank, world, local_rank).  Single process when WORLD_SIZE is unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def shard_indices(n_items: int, rank: int, world: int, weights=None):
    """Indices of the items rank `rank` processes.

    Default: strided split `rank::world` (the sorted path list of the reference, interleaved so that every rank
    sees the same mix).  With `weights` (e.g. estimated characters per line) items are assigned greedily,
    heaviest first, to the least loaded rank -- deterministic, independent of the calling rank."""
    if weights is None:
        return list(range(rank, n_items, world))
    order = sorted(range(n_items), key=lambda i: (-float(weights[i]), i))
    load = [0.0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += float(weights[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def broadcast_index(vectors: torch.Tensor | None, candidate_chars=None, src: int = 0):
    """Rank `src` passes the fp32 [N, D] prototype matrix (+ chars); every rank returns the same pair.
    One broadcast of the shape, one of the payload (77 MB for 50k x 384: far below a millisecond of NVLink time)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return vectors, candidate_chars
    rank = dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    meta = [None, None]
    if rank == src:
        meta = [tuple(vectors.shape), candidate_chars]
    dist.broadcast_object_list(meta, src=src)
    shape, chars = meta
    buf = vectors.to(dev, torch.float32).contiguous() if rank == src else torch.empty(shape, dtype=torch.float32, device=dev)
    dist.broadcast(buf, src=src)
    return buf, chars


def gather_results(local: dict, dst: int = 0):
    """{key: text} from every rank -> merged dict on rank `dst` (None elsewhere), keys sorted for determinism."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(sorted(local.items(), key=lambda kv: str(kv[0])))
    world = dist.get_world_size()
    out = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(local, out, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = {}
    for part in out:
        merged.update(part)
    return dict(sorted(merged.items(), key=lambda kv: str(kv[0])))
