"""Sample-index sharding of a progressive render across ranks (not in the reference, which is
one process on one GPU).

Samples are independent given (pixel, spp index, bias seed), but `tmpLifetime` consecutive samples share one
primary hit and one sub-pixel stratum (pathtracer.glsl:113-127, 206-211), so the unit of work is a whole
block k = spp [k*L, (k+1)*L). Block k goes to rank k mod world. Each rank adds its blocks' clamped radiance
into a SUM accumulator (adypt_tracer_accumulate); one all-reduce(sum) of W*H*4 floats combines the ranks and
adypt_tracer_resolve_sum divides by the sample count in .w. Batch ray tracing shards contiguous ray ranges
with no collective at all.
"""
from __future__ import annotations


def sample_blocks(total_spp: int, tmp_lifetime: int):
    """[(first_spp, n_spp)] covering [0, total_spp) in whole tmpLifetime blocks (the last may be short)."""
    out = []
    s = 0
    while s < total_spp:
        n = min(tmp_lifetime, total_spp - s)
        out.append((s, n))
        s += n
    return out


def blocks_for_rank(total_spp: int, tmp_lifetime: int, rank: int, world: int):
    return [b for k, b in enumerate(sample_blocks(total_spp, tmp_lifetime)) if k % world == rank]


def ray_range_for_rank(n: int, rank: int, world: int):
    """Contiguous [begin, end) of an n-ray batch for this rank (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def render_sharded(tracer, total_spp: int, rank: int, world: int, all_reduce_sum, preview_every: int = 0, preview=None):
    """Render this rank's blocks into the tracer's sum accumulator, reduce across ranks, resolve.

    tracer: object with .config.tmp_lifetime (or .tmp_lifetime), clear_sum(), accumulate(first, n), resolve_sum().
    all_reduce_sum: callable() that sums the tracer's accumulator over all ranks in place
    (torch.distributed.all_reduce on the tensor wrapping adypt_tracer_sum_buffer; a no-op when world == 1).
    preview_every / preview: progressive-preview cadence for interactive front ends. After every
    `preview_every` blocks of THIS rank, preview(blocks_done) is called on every rank (it is collective): the
    callee reduces a COPY of the accumulator (the accumulator itself keeps growing, so it must not be reduced in
    place) and shows sum.xyz / sum.w. Ranks with fewer blocks still join every preview so the collective matches.
    Returns the number of samples this rank rendered.
    """
    L = getattr(tracer, "tmp_lifetime", None) or tracer.config.tmp_lifetime
    tracer.clear_sum()
    mine = 0
    my_blocks = blocks_for_rank(total_spp, L, rank, world)
    most = max(len(blocks_for_rank(total_spp, L, r, world)) for r in range(world))
    for i in range(most):
        if i < len(my_blocks):
            first, n = my_blocks[i]
            tracer.accumulate(first, n)
            mine += n
        if preview is not None and preview_every > 0 and (i + 1) % preview_every == 0 and i + 1 < most:
            preview(i + 1)
    if world > 1:
        all_reduce_sum()
    tracer.resolve_sum()
    return mine


class DeviceArray:
    """Wraps a raw device pointer so torch.as_tensor(obj, device='cuda') can view it without copying."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
