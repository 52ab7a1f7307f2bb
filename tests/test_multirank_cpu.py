"""N > 1 host logic on CPU: two gloo ranks.  The renderer inside the ranks is the ORACLE (tests
may use it); what is under test is the sharding/reduction logic bench.py and the CLI share."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dsrt_b200 as d
    import oracle
    from conftest import POSE_FILE, cam12_of
    sharding = __import__("importlib").import_module("dsrt_b200.sharding")

    oracle.lib().or_set_threads(2)
    W, H, SPP, DEPTH = 48, 27, 6, 6
    poses = d.read_pose_file(POSE_FILE)
    tris, mats = d.synth_mesh(3000)
    sc = oracle.Scene(tris, mats)

    # --- frame-parallel: every frame rendered exactly once, no collective on the data path
    frames = list(range(90, 96))
    mine = sharding.frames_of_rank(len(frames), rank, world)
    sums = {}
    for k in mine:
        cam_pos, sun, _ = d.pose_to_model_frame(poses[frames[k]])
        cam12 = cam12_of(d.camera_look_at_origin(cam_pos, 40.0, W, H))
        acc, rgb, _ = sc.render(cam12, W, H, SPP, DEPTH, sun, rng_mode=0)
        sums[frames[k]] = float(acc.sum())
    gathered = [None] * world
    dist.all_gather_object(gathered, sums)          # control plane only: which rank did what
    # --- sample-sharded single frame: disjoint global sample ranges + ONE sum-reduce
    cam_pos, sun, _ = d.pose_to_model_frame(poses[95])
    cam12 = cam12_of(d.camera_look_at_origin(cam_pos, 40.0, W, H))
    s0, s1 = sharding.sample_range(SPP, rank, world)
    part, _, _ = sc.render(cam12, W, H, SPP, DEPTH, sun, rng_mode=0, s_begin=s0, s_end=s1)
    t = torch.from_numpy(part.copy())
    sharding.reduce_accum(t, dst=0)
    if rank == 0:
        full, rgb_full, _ = sc.render(cam12, W, H, SPP, DEPTH, sun, rng_mode=0)
        np.savez(Path(out_dir) / "result.npz", reduced=t.numpy(), full=full, rgb_full=rgb_full,
                 rgb_reduced=oracle.tonemap(t.numpy(), SPP, 2.0),
                 frames=np.array(sorted(f for g in gathered for f in g)), ranges=np.array([s0, s1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / "result.npz")
    assert r["frames"].tolist() == list(range(90, 96))                 # each frame exactly once
    # the shards add up to the single-rank frame (same global sample indices, A-6)
    assert np.allclose(r["reduced"], r["full"], atol=1e-5)
    assert (r["rgb_reduced"] == r["rgb_full"]).all(2).mean() > 0.999


def test_partition_rules():
    sharding = __import__("importlib").import_module("dsrt_b200.sharding")
    for world in (1, 2, 3, 4, 8):
        got = sorted(f for r in range(world) for f in sharding.frames_of_rank(99, r, world))
        assert got == list(range(99))
        sizes = [len(sharding.frames_of_rank(99, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
        for spp in (1, 7, 64, 250, 1000):
            cover = []
            for r in range(world):
                a, b = sharding.sample_range(spp, r, world)
                assert 0 <= a <= b <= spp
                cover += list(range(a, b))
            assert cover == list(range(spp))


def test_lpt_partition_of_the_pose_sequence():
    """bench.py --config c3 gives every rank a longest-processing-time-first share of the 99 poses: every pose exactly
    once, dearest first within a rank, the heaviest estimated load within 6 % of the mean, and never worse balanced
    than taking every N-th pose (the last poses of the approach cost several times the first ones)."""
    import dsrt_b200 as d
    from conftest import POSE_FILE
    sharding = __import__("importlib").import_module("dsrt_b200.sharding")
    poses = d.read_pose_file(POSE_FILE)
    costs = [sharding.frame_cost(d.pose_to_model_frame(p)[2]) for p in poses]
    assert costs[98] > 3 * costs[0]
    for world in (1, 2, 4, 8):
        parts = sharding.lpt_assignment(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(99))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) <= 1.06 * sum(costs) / world
        for p in parts:
            assert [costs[i] for i in p] == sorted((costs[i] for i in p), reverse=True)
        naive = [sum(costs[i] for i in sharding.frames_of_rank(99, r, world)) for r in range(world)]
        assert max(loads) <= max(naive) + 1e-9
    assert sharding.lpt_assignment(costs, 8) == sharding.lpt_assignment(costs, 8)      # the same table on every rank
