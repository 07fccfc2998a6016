"""Multi-GPU plumbing of the render path: one process per GPU, scene replicated, sample
passes sharded across ranks, accumulation buffers summed onto rank 0.

The reference shards by interleaved columns across web workers with the scene rebuilt in
every worker (src/worker.js:23-32, src/renderers.js:21,88) and composites on the main
thread (src/raytrace_launcher.js:92-97).  Passes are independent too (src/renderers.js:87),
and sharding them keeps every GPU on full frames (best occupancy) with one exchange:
a SUM reduce of the W x H float4 buffer (NCCL over NVLink on GPUs, gloo in CPU tests).
The RNG is keyed by the absolute pass index, so the image does not depend on the number
of ranks (up to float summation order).
"""
from __future__ import annotations


def pass_block(step: int, rank: int, world: int, passes_per_step: int):
    """(first_pass, n_passes) rendered by `rank` in `step`: blocks of `passes_per_step`
    consecutive pass indices dealt round-robin to the ranks."""
    first = (step * world + rank) * passes_per_step
    return first, passes_per_step


def shard_passes(total_passes: int, rank: int, world: int):
    """Strong-scaling split of `total_passes` pass indices: rank r gets {p : p % world == r}
    as a list of (first_pass, n_passes) runs of length 1."""
    return [(p, 1) for p in range(rank, total_passes, world)]


def reduce_accum(accum, dst: int = 0):
    """SUM-reduce the accumulation tensor onto `dst` (no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
