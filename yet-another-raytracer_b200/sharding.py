"""Multi-GPU sharding of the render: one process per GPU, the scene replicated,
the SAMPLE RANGE of every pixel split across ranks, and the per-rank f64 XYZ films summed onto
rank 0 with one collective (NCCL over NVLink on GPUs, gloo in the CPU tests).

Sample sharding is exact by construction: sanitize_sample_xyz acts per sample (reference
main.rs:700-707), so the film is a plain sum over samples and Philox counters make every sample
independent of who renders it.  The closest-hit sweep shards its ray array instead and needs no
collective at all.
"""


def shard_range(begin, end, rank, world):
    """Contiguous, balanced split of [begin, end) -- rank r gets the r-th piece (sizes differ by <= 1)."""
    n = max(0, end - begin)
    base, extra = divmod(n, world)
    lo = begin + rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def step_sample_range(step, rank, world, spp_per_step):
    """Weak-scaling schedule of bench.py: in every step each rank renders its own spp_per_step samples."""
    lo = (step * world + rank) * spp_per_step
    return lo, lo + spp_per_step


def reduce_film(film, dst=0, group=None):
    """Sum the per-rank films onto rank `dst` (in place).  `film` is a torch tensor on the device the
    process group was created for."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return film


def render_distributed(render_fn, film, sample_begin, sample_end, rank, world, dst=0, group=None):
    """Render this rank's share of [sample_begin, sample_end) into `film` (a zeroed torch tensor of shape
    (H, W, 3), f64) with render_fn(lo, hi, film) and combine on rank dst.  Returns this rank's (lo, hi)."""
    lo, hi = shard_range(sample_begin, sample_end, rank, world)
    if hi > lo:
        render_fn(lo, hi, film)
    reduce_film(film, dst=dst, group=group)
    return lo, hi
