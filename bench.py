#!/usr/bin/env python
"""bench.py -- throughput of the render hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode frame|sample]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one frame of BASELINE.json's config 2: pose 0 of the
shipped pose file, the synthetic 1M-triangle station, 1920x1080, 64 spp, depth 10, directional
sun with shadow rays.  Metric: Mrays/s (primary + bounce + shadow rays traced / device time).

  value      scene resident in HBM, CUDA events on the ctx stream around each step, L2
             flushed between steps, max over ranks
  e2e        the same frames through the C ABI with host buffers: camera/sun/params passed
             from host memory, tone-mapped RGB8 frame read back to host every step
  roofline   BVH traversal kernels (k_primary + k_trace share one traversal loop): algorithmic
             bytes B(N) x rays / their summed launch durations (CUDA events on the ctx stream),
             against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's CPU camera/bvh_node path (oracle/_ref/libref_cpu.so) on the
             host cores, on a bounded sample of the same frame (rank 0, N = 1 only)

N > 1: frame-parallel -- every rank renders its own frame of the same configuration (pose 0) per
step, no collective (scaling "weak": per-GPU work fixed, N frames per step).  --poses sequence
walks the pose file instead (rank r renders pose step*N + r; later poses are closer to the station
and cost more, so steps are then not comparable across N).  --mode sample shards ONE frame's samples over the
ranks and sums the float accumulation buffers with one NCCL reduce ("strong").
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
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
    ap.add_argument("--cpu_step", type=int, default=4, help="pixel stride of the CPU sample")
    ap.add_argument("--spb", type=int, default=0, help="samples of a pixel in flight together (0 = library default)")
    ap.add_argument("--config", default="", choices=["", "c2", "c4", "c5"],
                    help="BASELINE.json presets: c2 = the default; c4 = 3840x2160, 1000 spp (use with --mode sample); "
                         "c5 = 10M triangles, diffuse only, depth 16, 1080p, 256 spp")
    ap.add_argument("--no_ref_gpu", action="store_true", help="skip timing the reference's own GPU kernel beside the result")
    ap.add_argument("--poses", default="fixed", choices=["fixed", "sequence"],
                    help="frame-parallel runs: every rank renders pose 0 (fixed, the default) or rank r renders pose step*N + r")
    a = ap.parse_args()
    if a.config == "c4":
        a.width, a.height, a.spp = 3840, 2160, 1000
    elif a.config == "c5":
        a.tris, a.diffuse_only, a.depth, a.spp = 10_000_000, True, 16, 256
    return a


def workload_name(a):
    return (f"pose 0 of rendezvous_1s_dt0_01s.txt, synthetic {a.tris}-triangle station"
            f"{' (diffuse only)' if a.diffuse_only else ''}, {a.width}x{a.height}, {a.spp} spp, depth {a.depth}, "
            "directional sun + shadow rays")


# --------------------------------------------------------------------------------------------
def run_reference(a, rank, world):
    """The reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    import dsrt_b200 as d          # host helpers only (mesh, pose); no GPU call is made here
    import refcpu
    poses = d.read_pose_file(POSE_FILE)
    cam_pos, sun, _ = d.pose_to_model_frame(poses[0])
    tris, mats = d.synth_mesh(a.tris, diffuse_only=a.diffuse_only)
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
        import oracle
        kind = "port"
        t0 = time.perf_counter()
        sc = oracle.Scene(tris, mats)
        build_s = time.perf_counter() - t0
        cores = os.cpu_count() or 1
        cam = d.camera_look_at_origin(cam_pos, 40.0, a.width // step, a.height // step)
        cam12 = np.array(cam.origin.tuple() + cam.lower_left_corner.tuple() + cam.horizontal.tuple() + cam.vertical.tuple(), np.float32)

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
        "config": {"workload": workload_name(a), "sample": sample, "bvh_build_s": build_s},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------
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
    params = d.make_params(a.width, a.height, a.spp, a.depth, d.RNG_PHILOX)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    # the e2e leg reads every frame back into page-locked host memory (the caller's buffer; any host pointer works)
    rgb = torch.empty((a.height, a.width, 3), dtype=torch.uint8, pin_memory=True).numpy()

    def frame_inputs(i):
        # frame-parallel: step i covers frames i*world .. i*world+world-1, rank r takes its own
        idx = 0
        if world > 1 and a.mode == "frame" and a.poses == "sequence":
            idx = (sharding.frames_of_rank((i + 1) * world, rank, world)[i]) % len(poses)
        cam_pos, sun, _ = d.pose_to_model_frame(poses[idx])
        return d.camera_look_at_origin(cam_pos, 40.0, a.width, a.height, a.spp, a.depth), sun

    s_lo, s_hi = 0, a.spp
    if world > 1 and a.mode == "sample":
        s_lo, s_hi = sharding.sample_range(a.spp, rank, world)

    def reduce_accum():
        # one NCCL sum-reduce of the float accumulation buffer to rank 0 (NVLink/NVSwitch)
        ptr, n = ctx.accum_device()

        class _P:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        sharding.reduce_accum(torch.as_tensor(_P(), device=dev), dst=0)

    def step(i, timed):
        cam, sun = frame_inputs(i)
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fs = ctx.render(cam, sun, params, sample_begin=s_lo, sample_end=s_hi)
        if world > 1 and a.mode == "sample":
            with torch.cuda.stream(stream):
                reduce_accum()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1), fs

    for i in range(a.warmup):
        step(i, False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    clocks = ClockSampler(local_rank)
    clocks.start()
    ms = rays = launches = ext_ms = ext_rays = prim_ms = shade_ms = culled = 0.0
    for i in range(a.steps):
        m, fs = step(a.warmup + i, True)
        ms += m
        rays += fs.rays
        launches += fs.kernel_launches
        ext_ms += fs.extend_ms
        prim_ms += fs.primary_ms
        shade_ms += fs.shade_ms
        ext_rays += fs.rays
        culled += fs.culled_primary_rays
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    clk = clocks.stop()

    # ---- end to end through the C ABI with host buffers (render + tone-map + D2H each step)
    e2e_ms = e2e_rays = 0.0
    if not a.no_e2e:
        for i in range(a.steps):
            cam, sun = frame_inputs(a.warmup + i)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fs = ctx.render(cam, sun, params, sample_begin=s_lo, sample_end=s_hi)
            if world > 1 and a.mode == "sample":
                with torch.cuda.stream(stream):
                    reduce_accum()
                torch.cuda.synchronize(dev)
            if not (world > 1 and a.mode == "sample" and rank != 0):
                ctx.readback_rgb8(a.spp, params.gamma, rgb)
            e2e_ms += 1e3 * (time.perf_counter() - t0)
            e2e_rays += fs.rays

    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = t.tolist()
        s = torch.tensor([rays, e2e_rays, launches, ext_ms, ext_rays], dtype=torch.float64, device=dev)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        rays, e2e_rays, launches, ext_ms, ext_rays = s.tolist()
        ext_ms /= world            # average per-rank time of the traversal kernel

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        bpr = algorithmic_bytes_per_ray(a.tris)
        achieved = (ext_rays / world) * bpr / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        traffic = l2_read = None
        # DRAM bytes of the traversal launches of one step, from the committed `ncu --set full`
        # capture of this same workload (profiles/trace_ncu_summary.json); null for other workloads
        prof = ROOT / "profiles" / "trace_ncu_summary.json"
        if prof.exists():
            try:
                pj = json.loads(prof.read_text())
                if pj.get("workload") == workload_name(a):
                    traffic = pj.get("dram_bytes_per_step")
                    l2_read = pj.get("l2_to_sm_read_bytes_per_step")
            except Exception:  # noqa: BLE001
                traffic = None
        out = {
            "metric": "Mrays/s (primary+bounce+shadow)", "value": rays / (ms * 1e3), "unit": "Mrays/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": "strong" if (world > 1 and a.mode == "sample") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "parallelism": f"{a.mode}-parallel x{world}" if world > 1 else "single GPU",
                       "rng": "philox4x32-10", "l2": "flushed between steps (256 MiB write)",
                       "frames_per_s": world * a.steps / (ms * 1e-3) if a.mode == "frame" else a.steps / (ms * 1e-3),
                       # generator parameters and a SHA-256 of the vertex buffer go with every number
                       "mesh": {"generator": "dsrt_synth_mesh", "seed": 1337, "triangles": int(a.tris), "diffuse_only": bool(a.diffuse_only),
                                "vertices_sha256": mesh_sha},
                       "bvh": {"wide_nodes": st.num_wide_nodes, "levels": st.num_levels, "build_ms": st.build_ms,
                               "build_ms_first_call": st_first.build_ms, "upload_ms": st.upload_ms, "sah": st.sah_cost,
                               "builder": os.environ.get("DSRT_BVH_BUILDER", "ploc")}},
            "clocks": clk,
            "e2e": {"value": (e2e_rays / (e2e_ms * 1e3)) if e2e_ms else None, "unit": "Mrays/s",
                    "h2d_bytes_per_step": 104 + 48 + 24 + 8, "d2h_bytes_per_step": a.width * a.height * 3,
                    "ms_per_step": e2e_ms / a.steps if e2e_ms else None},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_primary + k_trace (BVH8 traversal, closest and any hit)", "achieved": achieved,
                         "peak": peak, "peak_source": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                         "bytes_per_ray": bpr, "rays_per_step": ext_rays / world / a.steps,
                         "kernel_ms_per_step": ext_ms / a.steps, "k_primary_ms_per_step": prim_ms / a.steps,
                         "k_shade_ms_per_step": shade_ms / a.steps, "traffic": traffic,
                         # of rays_per_step (rank 0's steps): camera rays whose 8x4-pixel tile provably sees no triangle
                         # and that ended as misses without a traversal; the rest were traversed
                         "culled_primary_rays_per_step": culled / a.steps,
                         "traversed_mrays_per_s": ((ext_rays / world - culled) / (ext_ms * 1e-3) / 1e6) if ext_ms else None,
                         # what actually moved, next to the algorithmic figure: DRAM and L2->SM bytes of the same launches
                         # (ncu, profiles/trace_ncu_summary.json) over the live kernel time
                         "traffic_gbs": (traffic / (ext_ms / a.steps * 1e-3) / 1e9) if traffic and ext_ms else None,
                         "l2_to_sm_read_gbs": (l2_read / (ext_ms / a.steps * 1e-3) / 1e9) if l2_read and ext_ms else None,
                         "note": "the 1M-triangle scene (57 MB of nodes and records) is L2-resident, so the traversal is bound by "
                                 "instruction issue / the ALU pipe, not by HBM (profiles/r01_be_trace_ncu.txt: issue 60-75 %, ALU 50-56 %); "
                                 "achieved counts SURVEY 8(d)'s algorithmic bytes, most of which the caches serve"},
        }
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(a, tris, mats, poses)
        if world == 1 and not a.no_ref_gpu and a.tris <= 2_000_000:
            ctx.close()
            out["ref_gpu_kernel"] = ref_gpu_kernel(a, tris, mats, poses, rays / a.steps)
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def ref_gpu_kernel(a, tris, mats, poses, rays_per_frame):
    """The reference's own CUDA path on the same B200 and frame (its CPU BVH build + its
    megakernel, oracle/_ref/libref_gpu.so): reported beside the result, not part of it.  The
    reference does not count rays; Mrays/s uses this path's ray count for the same frame."""
    try:
        import dsrt_b200 as d
        import refgpu
        if not refgpu.available():
            return {"value": None, "note": "oracle/_ref/libref_gpu.so not present"}
        cam_pos, sun, _ = d.pose_to_model_frame(poses[0])
        ref = refgpu.RefScene(tris, mats, cam_pos, sun, a.width, a.height, a.spp, a.depth)
        ms = [ref.render()[1] for _ in range(3)]
        build_ms = ref.build_ms
        ref.close()
        best = min(ms[1:])
        return {"value": rays_per_frame / (best * 1e3), "unit": "Mrays/s", "kernel_ms_per_frame": best,
                "scene_build_ms_per_frame": build_ms,
                "note": "render_kernel of the unmodified reference (one thread per pixel, LCG), second and third of 3 frames; "
                        "its build_gpu_scene (CPU BVH + uploads, redone for every frame by its main()) is timed separately"}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "note": repr(e)}


def cpu_baseline(a, tris, mats, poses):
    """The reference CPU path on a bounded sample of the same frame (rank 0, N = 1)."""
    import dsrt_b200 as d
    cam_pos, sun, _ = d.pose_to_model_frame(poses[0])
    step = max(1, a.cpu_step)
    cores = os.cpu_count() or 1
    ns = min(8, a.spp)
    sample = f"every {step}th pixel in x and y of the {a.width}x{a.height} frame, {ns} of {a.spp} samples, depth {a.depth}"
    try:
        import refcpu
        if refcpu.available():
            sc = refcpu.RefCpuScene(tris, mats)
            secs, c = sc.render(cam_pos, sun, a.width, a.height, a.spp, a.depth, step_x=step, step_y=step, s_begin=0, s_end=ns, threads=cores)
            return {"value": float(c.sum()) / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                    "sample": sample, "bvh_build_s": sc.build_s, "seconds": secs}
        import oracle
        sc = oracle.Scene(tris, mats)
        w, h = a.width // step, a.height // step
        cam = d.camera_look_at_origin(cam_pos, 40.0, w, h)
        cam12 = np.array(cam.origin.tuple() + cam.lower_left_corner.tuple() + cam.horizontal.tuple() + cam.vertical.tuple(), np.float32)
        t0 = time.perf_counter()
        _, _, c = sc.render(cam12, w, h, a.spp, a.depth, sun, rng_mode=1, s_begin=0, s_end=ns)
        secs = time.perf_counter() - t0
        return {"value": float(c.sum()) / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                "sample": f"{w}x{h} down-sampled frame, {ns} of {a.spp} samples, depth {a.depth}", "seconds": secs}
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
