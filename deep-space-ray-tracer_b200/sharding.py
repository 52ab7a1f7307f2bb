"""How the render path is split over GPUs.  Pure host logic, no device code.

* Pose sequences are frame-parallel: frame i goes to rank i mod N, no collective.
* One frame can be sample-sharded: rank r renders the GLOBAL sample indices
  [spp*r/N, spp*(r+1)/N) of every pixel (the jitter depends on s/spp, reference
  gpu_render.cu:995-996), the float accumulation buffers are summed with ONE reduce, and the
  root tone-maps.  host/main.cpp implements the same two rules in C++.
"""
from __future__ import annotations


def frames_of_rank(num_frames: int, rank: int, world: int) -> list[int]:
    return list(range(rank, num_frames, world))


def sample_range(spp: int, rank: int, world: int) -> tuple[int, int]:
    return spp * rank // world, spp * (rank + 1) // world


def reduce_accum(accum, dst: int = 0):
    """Sum-reduce a torch tensor holding an accumulation buffer to rank `dst` (NCCL on GPUs,
    gloo in the CPU tests).  Returns the tensor (valid on `dst`)."""
    import torch.distributed as dist
    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
