"""Batch sharding across GPUs: one process per GPU, no data-path collective except one allgather
of the decoded ids.

Replaces ``torch.nn.DataParallel`` of the reference (model.py:115), which replicates the module
and scatters / gathers tensors from one process every forward.  Crops are independent (eval-mode
BatchNorm, no cross-sample op), so rank r recognises the contiguous slice
``[r*ceil(B/N), ...)`` of the global batch with its own full weight replica, and the only exchange
is an ``all_gather`` of ``int32 ids [B/N, steps]`` (+ lengths) - 53 KB per rank at B=4096, N=8.
The transport is ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of the global batch owned by ``rank`` (ceil split; trailing ranks may be
    short or empty)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    per = (batch + world - 1) // world
    lo = min(rank * per, batch)
    return lo, min(lo + per, batch)


def shard_size(batch: int, world: int) -> int:
    return (batch + world - 1) // world


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment. Returns (rank, local_rank, world)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


class PendingGather:
    """Handle of an ids allgather in flight (``allgather_ids_async``).  The collective runs on the backend's own stream
    behind the work already enqueued on the caller's stream; ``wait()`` makes the caller's CURRENT stream (not the
    host) wait for it and returns the global ``(ids [batch, steps], lengths [batch])``.  Until then the caller's
    stream is free: the next step's forward overlaps the exchange."""

    def __init__(self, work, recv, batch, world, steps, ids, lengths):
        self._work, self._recv, self._batch, self._world, self._steps = work, recv, batch, world, steps
        self._local = (ids, lengths)

    def wait(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._work is None:
            return self._local
        self._work.wait()
        recv, batch, world, steps = self._recv, self._batch, self._world, self._steps
        per = shard_size(batch, world)
        if world * per != batch:   # drop the padding rows of short / empty trailing shards
            keep = torch.cat([torch.arange(r * per, r * per + (shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0]),
                                           device=recv.device) for r in range(world)])
            recv = recv.index_select(0, keep)
        return recv[:, :steps].contiguous(), recv[:, steps].contiguous()


def allgather_ids_async(ids: torch.Tensor, lengths: torch.Tensor, batch: int, group=None) -> PendingGather:
    """Start gathering per-rank ``ids [b_r, steps]`` / ``lengths [b_r]`` into the global ``[batch, steps]`` / ``[batch]``
    on every rank.  Shards are padded to the common ceil size so one fixed-size collective suffices; ids and lengths
    travel in ONE buffer (lengths as an extra column)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return PendingGather(None, None, batch, 1, ids.shape[1], ids, lengths)
    per = shard_size(batch, world)
    steps = ids.shape[1]
    send = torch.zeros((per, steps + 1), dtype=torch.int32, device=ids.device)
    n = ids.shape[0]
    send[:n, :steps] = ids
    send[:n, steps] = lengths
    recv = torch.empty((world * per, steps + 1), dtype=torch.int32, device=ids.device)
    work = dist.all_gather_into_tensor(recv, send, group=group, async_op=True)
    return PendingGather(work, recv, batch, world, steps, ids, lengths)


def allgather_ids(ids: torch.Tensor, lengths: torch.Tensor, batch: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Blocking form (on the caller's stream) of ``allgather_ids_async``."""
    return allgather_ids_async(ids, lengths, batch, group).wait()


def recognize_sharded(recognize: Callable[..., Dict[str, torch.Tensor]], image: torch.Tensor,
                      overlap: torch.Tensor, scene: torch.Tensor, ious: torch.Tensor,
                      rank: int, world: int, gather: bool = True, group=None) -> Dict[str, torch.Tensor]:
    """Recognise this rank's shard of the GLOBAL batch tensors and (optionally) gather the ids.

    ``recognize(image, overlap, scene, ious) -> dict(logits, ids, lengths)`` is ``Model.recognize``
    (or a stand-in in the CPU tests).  Logits stay local (42 MB at B=4096; nobody downstream of
    the decoder needs them on every rank)."""
    B = image.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    if hi > lo:
        out = recognize(image[lo:hi], overlap[lo:hi], scene[lo:hi], ious[lo:hi])
        ids, lengths, logits = out["ids"], out["lengths"], out.get("logits")
    else:
        probe = recognize(image[:1], overlap[:1], scene[:1], ious[:1])   # shapes / dtypes only
        ids, lengths, logits = probe["ids"][:0], probe["lengths"][:0], None
    res = {"local_ids": ids, "local_lengths": lengths, "local_logits": logits, "bounds": (lo, hi)}
    if gather:
        res["ids"], res["lengths"] = allgather_ids(ids, lengths, B, group)
    return res
