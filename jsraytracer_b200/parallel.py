"""Multi-GPU plumbing of the render path: one process per GPU, scene replicated, sample
passes sharded across ranks, accumulation buffers summed onto rank 0.

The reference shards by interleaved columns across web workers with the scene rebuilt in
every worker (src/worker.js:23-32, src/renderers.js:21,88) and composites on the main
thread (src/raytrace_launcher.js:92-97).  Passes are independent too (src/renderers.js:87),
and sharding them keeps every GPU on full frames (best occupancy) with one exchange:
a SUM reduce of the W x H float4 buffer (NCCL over NVLink on GPUs, gloo in CPU tests).
The RNG is keyed by the absolute pass index, so the image does not depend on the number
of ranks (up to float summation order).

The reference's own scheme is here too: `column_stripe` gives rank r the `x_offset = r,
x_delt = world` arguments of `renderer.render` (interleaved columns), and `gather_columns`
assembles the disjoint stripes on rank 0 with one gather (each rank sends only its W / world
columns) — the low-latency split of a single pass (strong scaling of one frame).
"""
from __future__ import annotations


def pass_block(step: int, rank: int, world: int, passes_per_step: int):
    """(first_pass, n_passes) rendered by `rank` in `step`: blocks of `passes_per_step`
    consecutive pass indices dealt round-robin to the ranks."""
    first = (step * world + rank) * passes_per_step
    return first, passes_per_step


def strong_share(passes_per_step: int, rank: int, world: int) -> int:
    """Strong scaling of a fixed frame: how many of a step's `passes_per_step` passes `rank` renders (the shares differ by
    at most one and add up to the step)."""
    return passes_per_step // world + (1 if rank < passes_per_step % world else 0)


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


def column_stripe(rank: int, world: int):
    """(x_offset, x_delt) of `rank`: the interleaved columns a web worker renders (src/worker.js:30-32,
    src/renderers.js:21,88)."""
    return rank, world


def gather_columns(accum, dst: int = 0):
    """Assemble column-striped buffers: `accum` is (H, W, C) and this rank rendered columns rank::world of it.
    Every rank sends its own columns only (padded to ceil(W / world)); `dst` interleaves them in place —
    the compositing of src/raytrace_launcher.js:92-97 as one gather.  No-op for a single process."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return accum
    rank, world = dist.get_rank(), dist.get_world_size()
    H, W, C = accum.shape
    ncols_max = (W + world - 1) // world
    send = torch.zeros((H, ncols_max, C), dtype=accum.dtype, device=accum.device)
    mine = accum[:, rank::world, :]
    send[:, : mine.shape[1], :] = mine
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank == dst:
        for r in range(world):
            if r != rank:
                n = len(range(r, W, world))
                accum[:, r::world, :] = recv[r][:, :n, :]
    return accum
