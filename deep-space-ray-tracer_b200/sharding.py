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


def frame_cost(sep_m: float) -> float:
    """Relative cost of one frame of the shipped kind of sequence: a floor plus the share of the image the station
    covers, which grows as the camera closes in and saturates when the station fills the frame (fitted to per-frame
    device times of the 99-pose sequence at 1080p; only the ratios matter)."""
    return 60.0 + 190.0 * min(1.0, (300.0 / max(sep_m, 1e-9)) ** 2)


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
