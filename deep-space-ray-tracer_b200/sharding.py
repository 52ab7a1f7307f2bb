"""How the render path is split over GPUs.  Pure host logic, no device code.

* Pose sequences are frame-parallel, no collective: frame i goes to rank i mod N (frames_of_rank), or -- when frames
  differ in cost, as in the shipped approach sequence -- ranks take a longest-processing-time-first partition
  (lpt_assignment over frame_cost).
* One frame can be sample-sharded: rank r renders the GLOBAL sample indices
  [spp*r/N, spp*(r+1)/N) of every pixel (the jitter depends on s/spp, reference
  gpu_render.cu:995-996), the float accumulation buffers are summed with ONE reduce, and the
  root tone-maps.  host/main.cpp implements the same two rules in C++.
"""
from __future__ import annotations


def frames_of_rank(num_frames: int, rank: int, world: int) -> list[int]:
    return list(range(rank, num_frames, world))


# Device time (ms) of one frame of the shipped approach sequence at seven camera distances (m), 1080p, 250 spp, depth 50,
# 1M triangles (profiles/r02_pose_split.txt, final code); only the ratios matter.
_COST_TABLE = ((35.7, 304.0), (71.5, 147.0), (143.0, 78.0), (268.0, 81.5), (536.1, 50.0), (1072.2, 54.0), (1787.0, 37.0))


def frame_cost(sep_m: float) -> float:
    """Relative cost of one frame of the shipped kind of sequence: the measured per-frame device times, interpolated
    in log(distance) -- flat while the station is a small part of the image, about twice that once it fills it, then
    growing like 1/distance as the camera moves into the structure and every path bounces more often."""
    import math
    t = _COST_TABLE
    x = math.log(min(max(sep_m, t[0][0]), t[-1][0]))
    for (s0, c0), (s1, c1) in zip(t, t[1:]):
        if x <= math.log(s1):
            w = (x - math.log(s0)) / (math.log(s1) - math.log(s0))
            return c0 + w * (c1 - c0)
    return t[-1][1]


def lpt_assignment(costs: list[float], world: int) -> list[list[int]]:
    """Longest-processing-time-first partition of frames over ranks: frames in decreasing cost, each to the rank
    with the least load so far (ties: lowest rank).  Deterministic, so every rank computes the same table.
    Returns per rank its frames, dearest first.  host/main.cpp gets the same effect with a shared counter over
    the cost-sorted list."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return out


def sample_range(spp: int, rank: int, world: int) -> tuple[int, int]:
    return spp * rank // world, spp * (rank + 1) // world


def reduce_accum(accum, dst: int = 0):
    """Sum-reduce a torch tensor holding an accumulation buffer to rank `dst` (NCCL on GPUs,
    gloo in the CPU tests).  Returns the tensor (valid on `dst`)."""
    import torch.distributed as dist
    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
