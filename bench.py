#!/usr/bin/env python
"""bench.py -- throughput of the render hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode frame|sample] [--config c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one frame of BASELINE.json's config 2: pose 0 of the
shipped pose file, the synthetic 1M-triangle station, 1920x1080, 64 spp, depth 10, directional
sun with shadow rays.  Metric: Mrays/s (primary + bounce + shadow rays traced / device time).

  value      scene resident in HBM, CUDA events on the ctx stream around each step, L2
             flushed between steps, max over ranks
  e2e        the same frames through the C ABI with host buffers: camera/sun/params passed
             from host memory, tone-mapped RGB8 frame copied to page-locked host memory every
             step, frames pipelined over two frame slots (dsrt_render_frame_async) the way the
             CLI drives a pose sequence; the one-frame-at-a-time figure is beside it
  roofline   BVH traversal kernels (k_primary + k_trace share one traversal loop): algorithmic
             bytes B(N) x TRAVERSED rays / their summed launch durations (CUDA events on the ctx
             stream), against MEASURED_PEAKS.json hbm_gbs; camera rays ended by tile culling
             without a traversal are not charged
  cpu_baseline  the reference's CPU camera/bvh_node path (oracle/_ref/libref_cpu.so) on the host
             cores on BASELINE config 1 (10k triangles, 400x225, 16 spp, depth 10, whole frame)
  c5         (N = 1) sub-record: BASELINE config 5 (10M triangles, diffuse, depth 16, 256 spp),
             the configuration that does not fit L2, with its own roofline
  c4_sample  (N > 1) sub-record: BASELINE config 4 (3840x2160, 1000 spp) sample-sharded over the
             ranks with one NCCL sum-reduce of the float accumulation buffer, reduce timed

N > 1: frame-parallel -- every rank renders its own frame of the same configuration (pose 0) per
step, no collective (scaling "weak": per-GPU work fixed, N frames per step).  --mode sample shards
ONE frame's samples over the ranks and sums the float accumulation buffers with one NCCL reduce
("strong").  --config c3 renders the 99-pose sequence (1080p, 250 spp, depth 50) frame-parallel with
PPM + PNG written and reports frames/s.

Only the baseline legs (cpu_baseline, ref_gpu_kernel, run_reference) touch oracle/; the timed product
path never does (tests/test_host.py checks this file's imports).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import queue
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

POSE_FILE = ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt"


def algorithmic_bytes_per_ray(ntris: int) -> float:
    """ray in (32 B) + hit out (16 B) + one root-to-leaf path through an 8-wide
    BVH of 80-byte nodes + a 4-triangle leaf of 48-byte records.  (The contract figure is fixed by the
    survey; the nodes this library builds are 96 bytes, csrc/bvh8.cuh.)"""
    levels = max(1, math.ceil(math.log(max(ntris, 8) / 4.0, 8)))
    return 48.0 + 80.0 * levels + 4 * 48.0


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dsrt", choices=["dsrt", "reference"])
    ap.add_argument("--mode", default="frame", choices=["frame", "sample"])
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--diffuse_only", action="store_true")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_e2e", action="store_true")
    ap.add_argument("--no_sub", action="store_true", help="skip the c5 (N = 1) / c4_sample (N > 1) sub-records")
    ap.add_argument("--cpu_step", type=int, default=4, help="pixel stride of the CPU sample of the reference arm")
    ap.add_argument("--spb", type=int, default=0, help="samples of a pixel in flight together (0 = library default)")
    ap.add_argument("--config", default="", choices=["", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json presets: c2 = the default; c3 = the 99-pose sequence at 1080p, 250 spp, depth 50, frames "
                         "written (frames/s); c4 = 3840x2160, 1000 spp (use with --mode sample); c5 = 10M triangles, diffuse only, "
                         "depth 16, 1080p, 256 spp")
    ap.add_argument("--no_ref_gpu", action="store_true", help="skip timing the reference's own GPU kernel beside the result")
    ap.add_argument("--no_write", action="store_true", help="c3: render and read back but do not write the frames")
    a = ap.parse_args()
    if a.config == "c4":
        a.width, a.height, a.spp = 3840, 2160, 1000
    elif a.config == "c5":
        a.tris, a.diffuse_only, a.depth, a.spp = 10_000_000, True, 16, 256
    elif a.config == "c3":
        a.width, a.height, a.spp, a.depth = 1920, 1080, 250, 50
    return a


def workload_name(tris, diffuse_only, width, height, spp, depth, what="pose 0"):
    return (f"{what} of rendezvous_1s_dt0_01s.txt, synthetic {tris}-triangle station"
            f"{' (diffuse only)' if diffuse_only else ''}, {width}x{height}, {spp} spp, depth {depth}, "
            "directional sun + shadow rays")


def env_overrides():
    """Tuning switches of the library that are set in the environment change what is measured: say so."""
    return {k: v for k, v in os.environ.items() if k.startswith("DSRT_")}


# --------------------------------------------------------------------------------------------
def run_reference(a, rank, world):
    """The reference's own CPU implementation of the path on the host cores.  Mesh, poses and camera
    come from oracle-side helpers: this arm never loads the product's library."""
    if rank != 0:
        return
    import oracle
    import refcpu
    import synthmesh
    poses = oracle.read_pose_file(POSE_FILE)
    cam_pos, sun, _ = oracle.pose_to_model_frame(poses[0])
    tris, mats = synthmesh.synth_mesh(a.tris, diffuse_only=a.diffuse_only)
    step = max(1, a.cpu_step)
    sample = (f"every {step}th pixel in x and y of the {a.width}x{a.height} frame, 1 of {a.spp} samples per step "
              f"(sample index = step), all {a.depth} bounces")
    if refcpu.available():
        kind = "reference"
        sc = refcpu.RefCpuScene(tris, mats)
        cores = os.cpu_count() or 1
        build_s = sc.build_s

        def one(i):
            s, c = sc.render(cam_pos, sun, a.width, a.height, a.spp, a.depth, step_x=step, step_y=step,
                             s_begin=i % a.spp, s_end=i % a.spp + 1, threads=cores)
            return s, int(c.sum())
    else:
        kind = "port"
        t0 = time.perf_counter()
        sc = oracle.Scene(tris, mats)
        build_s = time.perf_counter() - t0
        cores = os.cpu_count() or 1
        cam12, _ = oracle.camera_init(cam_pos, 40.0, a.width // step, a.height // step)

        def one(i):
            t0 = time.perf_counter()
            _, _, c = sc.render(cam12, a.width // step, a.height // step, a.spp, a.depth, sun, rng_mode=1,
                                s_begin=i % a.spp, s_end=i % a.spp + 1)
            return time.perf_counter() - t0, int(c.sum())
    for i in range(a.warmup):
        one(i)
    secs = rays = 0
    for i in range(a.steps):
        s, r = one(a.warmup + i)
        secs += s
        rays += r
    v = rays / secs / 1e6
    out = {
        "impl": "reference", "metric": "Mrays/s (primary+bounce+shadow)", "value": v, "unit": "Mrays/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * secs / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a.tris, a.diffuse_only, a.width, a.height, a.spp, a.depth), "sample": sample,
                   "bvh_build_s": build_s},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------
class Timed:
    """Sums of one timed loop."""

    def __init__(self):
        self.ms = self.rays = self.launches = self.ext_ms = self.prim_ms = self.shade_ms = self.culled = 0.0
        self.reduce_ms = 0.0


def roofline_record(tris, t: Timed, steps, world, traffic=None, l2_read=None):
    """achieved = algorithmic bytes of the rays that were TRAVERSED / the live time of the traversal kernels."""
    peak, peak_kind = measured_peak_gbs()
    bpr = algorithmic_bytes_per_ray(tris)
    rays = t.rays / world                      # per rank (ext_ms is the per-rank average)
    trav = rays - t.culled
    achieved = trav * bpr / (t.ext_ms * 1e-3) / 1e9 if t.ext_ms > 0 else 0.0
    all_rays = rays * bpr / (t.ext_ms * 1e-3) / 1e9 if t.ext_ms > 0 else 0.0
    in_l2 = tris <= 2_000_000
    note = ("the scene (57 MB of nodes and records at 1M triangles) is L2-resident: the traversal is bound by instruction issue and "
            "L1/L2 latency, not by HBM; `achieved` counts SURVEY 8(d)'s algorithmic bytes of the traversed rays, most of which the "
            "caches serve (`traffic` is what DRAM really moved)" if in_l2 else
            "the scene (about 0.6 GB of nodes and records at 10M triangles) does not fit L2: traversal is bound by DRAM/L2 latency "
            "under incoherent access; `traffic` is what DRAM really moved")
    return {"bound": "issue/L2" if in_l2 else "hbm", "kernel": "k_primary + k_trace (BVH8 traversal, closest and any hit)",
            "achieved": achieved, "peak": peak, "peak_source": peak_kind, "unit": "GB/s", "frac": achieved / peak,
            "frac_all_rays": all_rays / peak,
            "bytes_per_ray": bpr, "rays_per_step": rays / steps, "traversed_rays_per_step": trav / steps,
            # camera rays whose 8x4-pixel tile provably sees no triangle end as misses without a traversal; they count
            # in the headline Mrays/s (every path of the frame is resolved) but not here
            "culled_primary_rays_per_step": t.culled / steps,
            "traversed_mrays_per_s": (trav / (t.ext_ms * 1e-3) / 1e6) if t.ext_ms else None,
            "kernel_ms_per_step": t.ext_ms / steps, "k_primary_ms_per_step": t.prim_ms / steps,
            "k_shade_ms_per_step": t.shade_ms / steps, "traffic": traffic,
            "traffic_gbs": (traffic / (t.ext_ms / steps * 1e-3) / 1e9) if traffic and t.ext_ms else None,
            "traffic_frac_of_peak": (traffic / (t.ext_ms / steps * 1e-3) / 1e9 / peak) if traffic and t.ext_ms else None,
            "l2_to_sm_read_gbs": (l2_read / (t.ext_ms / steps * 1e-3) / 1e9) if l2_read and t.ext_ms else None,
            "note": note}


def committed_traffic(name, workload):
    """DRAM bytes of the traversal launches of one step from the committed `ncu --set full` capture of this same
    workload (profiles/<name>); null for other workloads."""
    prof = ROOT / "profiles" / name
    if prof.exists():
        try:
            pj = json.loads(prof.read_text())
            if pj.get("workload") == workload:
                return pj.get("dram_bytes_per_step"), pj.get("l2_to_sm_read_bytes_per_step")
        except Exception:  # noqa: BLE001
            pass
    return None, None


def run_dsrt(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import dsrt_b200 as d
    from dsrt_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    poses = d.read_pose_file(POSE_FILE)
    if a.config == "c3":
        run_sequence(a, rank, world, local_rank, torch, dist, d, sharding, dev, poses)
        return
    tris, mats = d.synth_mesh(a.tris, diffuse_only=a.diffuse_only)
    import hashlib
    mesh_sha = hashlib.sha256(np.ascontiguousarray(np.stack([np.stack([tris[v][k] for k in "xyz"], 1) for v in ("v0", "v1", "v2")], 1)).tobytes()).hexdigest()
    ctx = d.Context(local_rank)
    st_first = ctx.upload_scene(tris, mats)       # includes one-off module loading and allocator warm-up
    st = ctx.upload_scene(tris, mats)             # the same upload + GPU BVH build again: the repeatable cost
    ctx.set_stage_timing(True)
    if a.spb > 0:
        ctx.set_samples_per_batch(a.spb)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def frame_inputs(width, height, spp, depth):
        cam_pos, sun, _ = d.pose_to_model_frame(poses[0])
        return d.camera_look_at_origin(cam_pos, 40.0, width, height, spp, depth), sun

    def reduce_accum():
        # one NCCL sum-reduce of the float accumulation buffer to rank 0 (NVLink/NVSwitch)
        ptr, n = ctx.accum_device()

        class _P:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        sharding.reduce_accum(torch.as_tensor(_P(), device=dev), dst=0)

    def timed_loop(width, height, spp, depth, steps, warmup, sample_mode):
        """Device-timed steps: CUDA events on the ctx stream, L2 flushed before each, barrier + synchronize around."""
        params = d.make_params(width, height, spp, depth, d.RNG_PHILOX)
        cam, sun = frame_inputs(width, height, spp, depth)
        s_lo, s_hi = (sharding.sample_range(spp, rank, world) if (world > 1 and sample_mode) else (0, spp))
        t = Timed()

        def step(i):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize(dev)
            e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            fs = ctx.render(cam, sun, params, sample_begin=s_lo, sample_end=s_hi)
            em.record(stream)
            if world > 1 and sample_mode:
                with torch.cuda.stream(stream):
                    reduce_accum()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1), em.elapsed_time(e1), fs

        for i in range(warmup):
            step(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        clocks = ClockSampler(local_rank)
        clocks.start()
        for i in range(steps):
            m, r, fs = step(warmup + i)
            t.ms += m
            t.reduce_ms += r
            t.rays += fs.rays
            t.launches += fs.kernel_launches
            t.ext_ms += fs.extend_ms
            t.prim_ms += fs.primary_ms
            t.shade_ms += fs.shade_ms
            t.culled += fs.culled_primary_rays
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        clk = clocks.stop()
        if world > 1:
            mx = torch.tensor([t.ms, t.reduce_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            t.ms, t.reduce_ms = mx.tolist()
            sm = torch.tensor([t.rays, t.launches, t.ext_ms, t.prim_ms, t.shade_ms, t.culled], dtype=torch.float64, device=dev)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            t.rays, t.launches, t.ext_ms, t.prim_ms, t.shade_ms, t.culled = sm.tolist()
            t.ext_ms /= world            # average per-rank time of the traversal kernels
            t.prim_ms /= world
            t.shade_ms /= world
            t.culled /= world
        return t, clk, params, cam, sun, (s_lo, s_hi)

    sample_mode = a.mode == "sample"
    t, clk, params, cam, sun, (s_lo, s_hi) = timed_loop(a.width, a.height, a.spp, a.depth, a.steps, a.warmup, sample_mode)

    # ---- end to end through the C ABI with host buffers (render + tone-map + D2H each step)
    e2e_ms = e2e_rays = e2e_sync_ms = 0.0
    if not a.no_e2e:
        rgbs = [torch.empty((a.height, a.width, 3), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
        # (1) one frame at a time: render, then tone map + copy, each waited for
        for i in range(a.steps):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fs = ctx.render(cam, sun, params, sample_begin=s_lo, sample_end=s_hi)
            if world > 1 and sample_mode:
                with torch.cuda.stream(stream):
                    reduce_accum()
                torch.cuda.synchronize(dev)
            if not (world > 1 and sample_mode and rank != 0):
                ctx.readback_rgb8(a.spp, params.gamma, rgbs[0])
            e2e_sync_ms += 1e3 * (time.perf_counter() - t0)
        # (2) the way a pose sequence is driven: two frames in flight on two frame slots, every frame's camera and sun
        # passed from host memory and its RGB8 image copied to page-locked host memory inside the timed region
        if world > 1 and sample_mode:
            e2e_ms, e2e_rays = e2e_sync_ms, t.rays / max(world, 1)
        else:
            for k in range(2):             # warm both slots (work memory of slot 1 is allocated at first use)
                ctx.render_async(k, cam, sun, params)
                ctx.readback_rgb8_async(k, a.spp, params.gamma, rgbs[k])
            for k in range(2):
                ctx.frame_wait(k)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for i in range(a.steps):
                k = i & 1
                if i >= 2:
                    e2e_rays += ctx.frame_wait(k).rays
                cam_i, sun_i = frame_inputs(a.width, a.height, a.spp, a.depth)        # host-side pose transform + camera, per frame
                ctx.render_async(k, cam_i, sun_i, params)
                ctx.readback_rgb8_async(k, a.spp, params.gamma, rgbs[k])
            for i in range(max(0, a.steps - 2), a.steps):
                e2e_rays += ctx.frame_wait(i & 1).rays
            e2e_ms = 1e3 * (time.perf_counter() - t0)

    if world > 1:
        mx = torch.tensor([e2e_ms, e2e_sync_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_sync_ms = mx.tolist()
        sm = torch.tensor([e2e_rays], dtype=torch.float64, device=dev)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_rays = sm.tolist()[0]

    # ---- sub-records
    sub = {}
    if not a.no_sub and a.config in ("", "c2"):
        if world > 1:
            # BASELINE config 4: one 3840x2160 frame at 1000 spp, samples sharded over the ranks, one NCCL reduce
            t4, clk4, _, _, _, _ = timed_loop(3840, 2160, 1000, a.depth, 3, 1, True)
            # the reduce alone, all ranks entering together (the figure inside the step also holds the wait for the slowest rank)
            alone = []
            for _ in range(3):
                dist.barrier()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                with torch.cuda.stream(stream):
                    reduce_accum()
                e1.record(stream)
                torch.cuda.synchronize(dev)
                alone.append(e0.elapsed_time(e1))
            ar = torch.tensor([min(alone)], dtype=torch.float64, device=dev)
            dist.all_reduce(ar, op=dist.ReduceOp.MAX)
            sub["c4_sample"] = {
                "workload": workload_name(a.tris, a.diffuse_only, 3840, 2160, 1000, a.depth),
                "parallelism": f"sample-parallel x{world}: rank r renders global samples [1000 r/{world}, 1000 (r+1)/{world}), "
                               "one NCCL sum-reduce of the float accumulation buffer to rank 0",
                "steps": 3, "warmup": 1, "ms_per_frame": t4.ms / 3, "value": t4.rays / (t4.ms * 1e3), "unit": "Mrays/s",
                "scaling": "strong", "reduce_bytes": 3840 * 2160 * 3 * 4,
                "reduce_ms_in_step": t4.reduce_ms / 3, "reduce_ms_alone": ar.tolist()[0],
                "reduce_alone_gbs": 3840 * 2160 * 3 * 4 / (ar.tolist()[0] * 1e-3) / 1e9,
                "clocks": clk4}
    if rank == 0:
        wl = workload_name(a.tris, a.diffuse_only, a.width, a.height, a.spp, a.depth)
        traffic, l2_read = committed_traffic("trace_ncu_summary.json", wl)
        if traffic is None:
            traffic, l2_read = committed_traffic("trace_ncu_summary_c5.json", wl)
        out = {
            "metric": "Mrays/s (primary+bounce+shadow)", "value": t.rays / (t.ms * 1e3), "unit": "Mrays/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": t.ms / a.steps,
            "higher_is_better": True, "scaling": "strong" if (world > 1 and sample_mode) else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "parallelism": f"{a.mode}-parallel x{world}" if world > 1 else "single GPU",
                       "rng": "philox4x32-10", "l2": "flushed between steps (256 MiB write)",
                       "frames_per_s": world * a.steps / (t.ms * 1e-3) if a.mode == "frame" else a.steps / (t.ms * 1e-3),
                       # generator parameters and a SHA-256 of the vertex buffer go with every number
                       "mesh": {"generator": "dsrt_synth_mesh", "seed": 1337, "triangles": int(a.tris), "diffuse_only": bool(a.diffuse_only),
                                "vertices_sha256": mesh_sha},
                       "bvh": {"wide_nodes": st.num_wide_nodes, "levels": st.num_levels, "build_ms": st.build_ms,
                               "build_ms_first_call": st_first.build_ms, "upload_ms": st.upload_ms, "sah": st.sah_cost,
                               "builder": os.environ.get("DSRT_BVH_BUILDER", "ploc")},
                       "env_overrides": env_overrides()},
            "clocks": clk,
            "e2e": {"value": (e2e_rays / (e2e_ms * 1e3)) if e2e_ms else None, "unit": "Mrays/s",
                    "h2d_bytes_per_step": 104 + 48 + 24 + 8, "d2h_bytes_per_step": a.width * a.height * 3,
                    "ms_per_step": e2e_ms / a.steps if e2e_ms else None,
                    "how": "two frames in flight on two frame slots (dsrt_render_frame_async + dsrt_readback_rgb8_async into page-locked "
                           "host buffers + dsrt_frame_wait), wall clock over all steps",
                    "ms_per_step_one_at_a_time": e2e_sync_ms / a.steps if e2e_sync_ms else None},
            "gpu_launches": int(t.launches),
            "roofline": roofline_record(a.tris, t, a.steps, world, traffic, l2_read),
        }
        if world > 1 and sample_mode:
            out["reduce_ms_per_step"] = t.reduce_ms / a.steps
        out.update(sub)
    ctx.close()
    del flush
    if rank == 0:
        if world == 1 and not a.no_sub and a.config in ("", "c2"):
            out["c5"] = c5_record(torch, dev, d)
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline()
        if world == 1 and not a.no_ref_gpu and a.tris <= 2_000_000:
            out["ref_gpu_kernel"] = ref_gpu_kernel(a, t.rays / a.steps, (t.rays - t.culled) / a.steps)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c5_record(torch, dev, d):
    """BASELINE config 5 on this GPU: 10M triangles, diffuse only, depth 16, 1080p, 256 spp -- the scene that does not
    fit the 126 MB L2 (about 0.6 GB of nodes and triangle records)."""
    tris_n, w, h, spp, depth, steps, warmup = 10_000_000, 1920, 1080, 256, 16, 5, 2
    try:
        poses = d.read_pose_file(POSE_FILE)
        cam_pos, sun, _ = d.pose_to_model_frame(poses[0])
        cam = d.camera_look_at_origin(cam_pos, 40.0, w, h, spp, depth)
        tris, mats = d.synth_mesh(tris_n, diffuse_only=True)
        ctx = d.Context(dev.index)
        st = ctx.upload_scene(tris, mats)
        del tris
        ctx.set_stage_timing(True)
        stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        params = d.make_params(w, h, spp, depth, d.RNG_PHILOX)
        t = Timed()
        for i in range(warmup + steps):
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fs = ctx.render(cam, sun, params)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            if i < warmup:
                continue
            t.ms += e0.elapsed_time(e1)
            t.rays += fs.rays
            t.launches += fs.kernel_launches
            t.ext_ms += fs.extend_ms
            t.prim_ms += fs.primary_ms
            t.shade_ms += fs.shade_ms
            t.culled += fs.culled_primary_rays
        ctx.close()
        wl = workload_name(tris_n, True, w, h, spp, depth)
        traffic, l2_read = committed_traffic("trace_ncu_summary_c5.json", wl)
        return {"workload": wl, "steps": steps, "warmup": warmup, "ms_per_step": t.ms / steps, "value": t.rays / (t.ms * 1e3),
                "unit": "Mrays/s", "l2": "not flushed: the scene and the 19 GB of work memory of a step are far larger than L2",
                "bvh": {"wide_nodes": st.num_wide_nodes, "levels": st.num_levels, "build_ms": st.build_ms, "sah": st.sah_cost},
                "gpu_launches": int(t.launches), "roofline": roofline_record(tris_n, t, steps, 1, traffic, l2_read)}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def run_sequence(a, rank, world, local_rank, torch, dist, d, sharding, dev, poses):
    """BASELINE config 3: the whole pose sequence (99 poses), 1080p, 250 spp, depth 50 (the reference's defaults),
    frame-parallel: rank r renders poses r, r + N, ... with two frames in flight, PPM + PNG of every frame written by
    writer threads (what the reference's loop does per frame, src/main.cpp:310-431).  No collective.  The metric is
    frames/s over the wall clock of the frame loop (scene upload and CUDA start-up are reported beside it)."""
    t_start = time.perf_counter()
    tris, mats = d.synth_mesh(a.tris, diffuse_only=a.diffuse_only)
    ctx = d.Context(local_rank)
    st = ctx.upload_scene(tris, mats)
    params = d.make_params(a.width, a.height, a.spp, a.depth, d.RNG_PHILOX)
    # the camera closes in over the sequence and late poses cost ~6x the early ones: ranks take a longest-processing-time
    # partition of the poses (the same table on every rank, no communication), dearest frames first
    costs = [sharding.frame_cost(d.pose_to_model_frame(p)[2]) for p in poses]
    mine = sharding.lpt_assignment(costs, world)[rank]
    outdir = Path(tempfile.mkdtemp(prefix=f"dsrt_c3_r{rank}_"))
    bufs = queue.Queue()
    for _ in range(4):
        bufs.put(torch.empty((a.height, a.width, 3), dtype=torch.uint8, pin_memory=True).numpy())
    jobs = queue.Queue()
    written = []

    def writer():
        while True:
            it = jobs.get()
            if it is None:
                return
            idx, buf = it
            if not a.no_write:
                d.write_ppm(outdir / f"frame_{idx:04d}.ppm", buf, a.width, a.height)
                d.write_png(outdir / f"frame_{idx:04d}.png", buf, a.width, a.height)
            written.append(idx)
            bufs.put(buf)

    writers = [threading.Thread(target=writer) for _ in range(2)]
    for w in writers:
        w.start()

    def enqueue(slot, idx):
        cam_pos, sun, _ = d.pose_to_model_frame(poses[idx])
        cam = d.camera_look_at_origin(cam_pos, 40.0, a.width, a.height, a.spp, a.depth)
        buf = bufs.get()
        ctx.render_async(slot, cam, sun, params)
        ctx.readback_rgb8_async(slot, a.spp, params.gamma, buf)
        return idx, buf

    # warm-up: allocate both slots' work memory on the most expensive pose, outside the timed loop
    for k in range(2):
        _, buf = enqueue(k, len(poses) - 1)
        ctx.frame_wait(k)
        bufs.put(buf)
    setup_s = time.perf_counter() - t_start
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    inflight = [None, None]
    rays = dev_ms = 0.0
    for n, idx in enumerate(mine):
        k = n & 1
        if inflight[k] is not None:
            fs = ctx.frame_wait(k)
            rays += fs.rays
            dev_ms += fs.device_ms
            jobs.put(inflight[k])
        inflight[k] = enqueue(k, idx)
    for k in ((len(mine) & 1), (len(mine) & 1) ^ 1):
        if inflight[k] is not None:
            fs = ctx.frame_wait(k)
            rays += fs.rays
            dev_ms += fs.device_ms
            jobs.put(inflight[k])
    for _ in writers:
        jobs.put(None)
    for w in writers:
        w.join()
    loop_s = time.perf_counter() - t0
    ctx.close()
    nbytes = sum(f.stat().st_size for f in outdir.iterdir())
    for f in outdir.iterdir():
        f.unlink()
    outdir.rmdir()
    if world > 1:
        mx = torch.tensor([loop_s, setup_s], dtype=torch.float64, device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        loop_s, setup_s = mx.tolist()
        sm = torch.tensor([rays, dev_ms, float(len(written)), float(nbytes)], dtype=torch.float64, device=dev)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        rays, dev_ms, nwritten, nbytes = sm.tolist()
    else:
        nwritten = len(written)
    if rank == 0:
        out = {"metric": "frames/s (99-pose sequence, 1080p, 250 spp, depth 50, PPM + PNG written)", "value": nwritten / loop_s,
               "unit": "frames/s", "n_gpus": world, "steps": int(nwritten), "warmup": 2, "ms_per_step": 1e3 * loop_s / max(nwritten, 1),
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload_name(a.tris, a.diffuse_only, a.width, a.height, a.spp, a.depth, "all 99 poses"),
                          "parallelism": f"frame-parallel x{world}, longest-processing-time partition of the poses over the ranks; two frames in flight per GPU; no collective",
                          "frames_written": int(nwritten), "bytes_written": int(nbytes), "write": not a.no_write,
                          "env_overrides": env_overrides()},
               "sequence_wall_s": loop_s, "setup_s_max_over_ranks": setup_s, "sum_device_ms": dev_ms,
               "mrays_per_s": rays / (loop_s * 1e6), "bvh_build_ms": st.build_ms,
               "e2e": {"value": nwritten / loop_s, "unit": "frames/s", "h2d_bytes_per_step": 104 + 48 + 24 + 8,
                       "d2h_bytes_per_step": a.width * a.height * 3},
               "gpu_launches": None}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ref_gpu_kernel(a, rays_per_frame, traversed_per_frame):
    """The reference's own CUDA path on the same B200 and frame (its CPU BVH build + its
    megakernel, oracle/_ref/libref_gpu.so): reported beside the result, not part of it.  The
    reference does not count rays; Mrays/s uses this path's ray count for the same frame."""
    try:
        import oracle
        import refgpu
        import synthmesh
        if not refgpu.available():
            return {"value": None, "note": "oracle/_ref/libref_gpu.so not present"}
        poses = oracle.read_pose_file(POSE_FILE)
        cam_pos, sun, _ = oracle.pose_to_model_frame(poses[0])
        tris, mats = synthmesh.synth_mesh(a.tris, diffuse_only=a.diffuse_only)
        ref = refgpu.RefScene(tris, mats, cam_pos, sun, a.width, a.height, a.spp, a.depth)
        ms = [ref.render()[1] for _ in range(3)]
        build_ms = ref.build_ms
        ref.close()
        best = min(ms[1:])
        return {"value": rays_per_frame / (best * 1e3), "unit": "Mrays/s", "kernel_ms_per_frame": best,
                "traversed_mrays_per_s": traversed_per_frame / (best * 1e3),
                "scene_build_ms_per_frame": build_ms,
                "note": "render_kernel of the unmodified reference (one thread per pixel, LCG), second and third of 3 frames; "
                        "its build_gpu_scene (CPU BVH + uploads, redone for every frame by its main()) is timed separately; "
                        "the apples-to-apples figure is the frame time ratio"}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "note": repr(e)}


def cpu_baseline():
    """The reference CPU path on BASELINE config 1: pose 0, the synthetic 10k-triangle
    station, 400x225, 16 spp, depth 10, the WHOLE frame, all host cores."""
    import oracle
    import synthmesh
    tris_n, w, h, spp, depth = 10_000, 400, 225, 16, 10
    cores = os.cpu_count() or 1
    sample = f"BASELINE config 1, whole frame: {tris_n} triangles, {w}x{h}, {spp} spp, depth {depth}, pose 0"
    try:
        poses = oracle.read_pose_file(POSE_FILE)
        cam_pos, sun, _ = oracle.pose_to_model_frame(poses[0])
        tris, mats = synthmesh.synth_mesh(tris_n)
        import refcpu
        if refcpu.available():
            sc = refcpu.RefCpuScene(tris, mats)
            sc.render(cam_pos, sun, w, h, spp, depth, s_begin=0, s_end=1, threads=cores)        # warm the thread pool
            secs, c = sc.render(cam_pos, sun, w, h, spp, depth, threads=cores)
            return {"value": float(c.sum()) / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                    "sample": sample, "bvh_build_s": sc.build_s, "seconds_per_frame": secs, "rays": int(c.sum())}
        sc = oracle.Scene(tris, mats)
        cam12, _ = oracle.camera_init(cam_pos, 40.0, w, h)
        t0 = time.perf_counter()
        _, _, c = sc.render(cam12, w, h, spp, depth, sun, rng_mode=1)
        secs = time.perf_counter() - t0
        return {"value": float(c.sum()) / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                "sample": sample, "seconds_per_frame": secs, "rays": int(c.sum())}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "Mrays/s", "cores": cores, "kind": "unavailable", "sample": repr(e)}


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
    else:
        run_dsrt(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
