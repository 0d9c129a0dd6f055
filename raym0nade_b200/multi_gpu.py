"""Multi-GPU sharding of one render: interleaved sample indices + one reduction per frame.

Samples are i.i.d. given the (sample-independent) primary hit, so rank r of W takes the direct
and indirect sample indices s = r, r+W, r+2W, ... with the GLOBAL 1/spp weights and no
data-path collective while sampling.  At the end of the frame the fp32 accumulators are
combined in three steps (include/raym0nade_b200.h, "Multi-GPU reduction"):

  1. firefly-clamp side data (src/render.cpp:534-547): per pixel {sum of sample luminances,
     sample count} is SUM-reduced and the luminance of each rank's held-back sample is
     MAX-reduced;
  2. every rank commits its held-back sample against the global totals - only the rank that
     owns the global maximum can drop it (at most one sample per pixel can exceed 16/17 of
     the pixel's total);
  3. the 16 radiance floats per pixel are SUM-reduced to rank 0, which resolves.

`reduce_frame` is written against a small protocol (accum_view / accum_after_reduce /
accum_radiance returning array-likes) so the same code drives the CUDA context with NCCL and
a numpy stand-in with gloo in the CPU tests.
"""
from __future__ import annotations


def local_sample_count(total: int, rank: int, world: int) -> int:
    """how many of the sample indices 0..total-1 fall to `rank` (s = rank + k*world)"""
    return (total - rank + world - 1) // world if total > rank else 0


def reduce_frame(acc, dist, rank: int, world: int, as_tensor, dst: int = 0):
    """acc: object with accum_view() -> (sum_buf, max_buf), accum_after_reduce(rank, world),
    accum_radiance() -> rad_buf; as_tensor: buffer -> torch tensor sharing its memory."""
    if world <= 1:
        return
    sum_buf, max_buf = acc.accum_view()
    dist.all_reduce(as_tensor(sum_buf), op=dist.ReduceOp.SUM)
    dist.all_reduce(as_tensor(max_buf), op=dist.ReduceOp.MAX)
    acc.accum_after_reduce(rank, world)
    dist.reduce(as_tensor(acc.accum_radiance()), dst=dst, op=dist.ReduceOp.SUM)


def frame_slice(npix: int, rank: int, world: int):
    """(first pixel, pixel count) of the frame rank `rank` owns after the scattered exchange: slices of
    per = ceil(npix / world) pixels, the last ones clipped to the frame (rm_frame_slice)"""
    per = (npix + world - 1) // world
    first = min(rank * per, npix)
    return first, max(0, min(per, npix - first))


def reduce_frame_scatter(acc, dist, rank: int, world: int, as_tensor, npix: int):
    """The same exchange with its last step scattered (rm_reduce_scatter): afterwards the radiance buffer of rank r holds
    the frame's sums in ITS slice of the pixels only; returns that slice.  With a backend that has no reduce-scatter
    (gloo) every slice is reduced to its owner."""
    if world <= 1:
        return 0, npix
    sum_buf, max_buf = acc.accum_view()
    dist.all_reduce(as_tensor(sum_buf), op=dist.ReduceOp.SUM)
    dist.all_reduce(as_tensor(max_buf), op=dist.ReduceOp.MAX)
    acc.accum_after_reduce(rank, world)
    rad = as_tensor(acc.accum_radiance()).view(-1)
    for owner in range(world):
        first, count = frame_slice(npix, owner, world)
        if count:
            dist.reduce(rad[first * 16:(first + count) * 16], dst=owner, op=dist.ReduceOp.SUM)
    return frame_slice(npix, rank, world)


class ContextAccum:
    """adapter: raym0nade_b200.api.Context -> the protocol above (device pointers + lengths)"""

    def __init__(self, ctx):
        self.ctx = ctx

    def accum_view(self):
        ps, ns, pm, nm = self.ctx.accum_view()
        return (ps, ns), (pm, nm)

    def accum_after_reduce(self, rank, world):
        self.ctx.accum_after_reduce(rank, world)

    def accum_radiance(self):
        return self.ctx.accum_radiance()


class DevPtr:
    """__cuda_array_interface__ view of a raw fp32 device buffer"""

    def __init__(self, ptr_n):
        ptr, n = ptr_n
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
