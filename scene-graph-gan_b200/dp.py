"""Host-side data-parallel plumbing (the reference is single-GPU, train.py:417-418).

The path shards over the batch: rank r owns samples [r*B_local, (r+1)*B_local) of the global batch, every
loss is normalised by the GLOBAL batch inside the kernels (sgg_step_args_t.world), so summing the per-rank
gradient buckets reproduces the single-process gradient of the concatenated batch.  Nothing here touches
CUDA; tests/test_data_parallel_cpu.py runs it with the gloo backend at world_size 2.
"""
from __future__ import annotations

from typing import Callable, Tuple


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Equal shards only: the kernels normalise by B_local * world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by the world size {world}")
    b = global_batch // world
    return rank * b, (rank + 1) * b


def rank_seed(seed: int, rank: int) -> int:
    """Philox key of a rank: identical initial weights everywhere (same `seed`), decorrelated noise /
    interpolation-coefficient streams per rank."""
    return (seed * 1000003 + 7919 * rank) & 0x7FFFFFFF


def broadcast_comm_id(dist, group, rank: int, make_id: Callable[[], bytes]) -> bytes:
    """Rank 0 creates the communicator's unique id (sgg_comm_unique_id); everyone receives it."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    if not isinstance(box[0], (bytes, bytearray)) or len(box[0]) == 0:
        raise RuntimeError("communicator id was not delivered")
    return bytes(box[0])
