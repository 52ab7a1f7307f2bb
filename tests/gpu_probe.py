"""GPU bring-up probe (run on a B200 through gpurun; not a pytest file).

Runs the B200 path next to the unmodified reference renderer (oracle/_ref/libref_gpu.so) and
records what agrees and what does not in gpurun_out/probe.json.  Stages are independent so
one failure does not hide the others.

    python tests/gpu_probe.py [--stages bvh,first,rays,lcg,philox,perf] [--tris 10000]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import dsrt_b200 as d  # noqa: E402
import refgpu  # noqa: E402
from bvhcheck import check_bvh  # noqa: E402

POSE_FILE = ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt"
OUT = ROOT / "gpurun_out"


def pose0():
    poses = d.read_pose_file(POSE_FILE)
    cam, sun, sep = d.pose_to_model_frame(poses[0])
    return cam, sun


def random_rays(n, seed, lo, hi):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo - 50, hi + 50, size=(n, 3)).astype(np.float32)
    tgt = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    dirs = (tgt - o).astype(np.float32)
    # half the rays get unit length, half stay un-normalised like the reference's primaries
    ln = np.linalg.norm(dirs, axis=1, keepdims=True).astype(np.float32)
    dirs[: n // 2] = dirs[: n // 2] / ln[: n // 2]
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o
    rays[:, 3] = 0.001
    rays[:, 4:7] = dirs
    rays[:, 7] = 1e9
    return rays


def compare_hits(ours, ref):
    tri, t, u, v = ours
    rtri, rt, ru, rv = ref
    same_tri = tri == rtri
    hit = rtri >= 0
    t_bits = t.view(np.uint32) == rt.view(np.uint32)
    u_bits = u.view(np.uint32) == ru.view(np.uint32)
    v_bits = v.view(np.uint32) == rv.view(np.uint32)
    mism = np.nonzero(~same_tri)[0]
    ties = int(np.sum((~same_tri) & (tri >= 0) & (rtri >= 0) & t_bits))
    ref_missed = int(np.sum((~same_tri) & (rtri < 0) & (tri >= 0)))
    we_missed = int(np.sum((~same_tri) & (rtri >= 0) & (tri < 0)))
    both = same_tri & hit
    rel = np.abs(t[both] - rt[both]) / np.maximum(np.abs(rt[both]), 1e-30) if both.any() else np.zeros(1)
    return dict(n=int(len(tri)), ref_hits=int(hit.sum()), our_hits=int((tri >= 0).sum()),
                tri_equal=int(same_tri.sum()), tri_mismatch=int(len(mism)),
                mismatch_exact_t_ties=ties, mismatch_ref_missed=ref_missed, mismatch_we_missed=we_missed,
                t_bits_equal_on_same_tri=int((t_bits & both).sum()), u_bits_equal=int((u_bits & both).sum()),
                v_bits_equal=int((v_bits & both).sum()), same_tri_hits=int(both.sum()),
                max_rel_t=float(rel.max()), first_mismatches=[int(x) for x in mism[:8]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", default="bvh,first,rays,lcg,philox,perf")
    ap.add_argument("--tris", type=int, default=10000)
    ap.add_argument("--perf_tris", type=int, default=1000000)
    ap.add_argument("--out", default=str(OUT / "probe.json"))
    args = ap.parse_args()
    stages = args.stages.split(",")
    OUT.mkdir(exist_ok=True)
    res = {"version": d.version(), "ref_gpu": refgpu.available()}

    W, H, SPP, DEPTH = 400, 225, 16, 10
    cam_pos, sun = pose0()
    tris, mats = d.synth_mesh(args.tris)
    ctx = d.Context(0)
    st = ctx.upload_scene(tris, mats)
    res["scene"] = dict(tris=st.num_triangles, nodes=st.num_wide_nodes, levels=st.num_levels,
                        build_ms=st.build_ms, upload_ms=st.upload_ms, sah=st.sah_cost)
    print("scene", res["scene"], flush=True)
    cam = d.camera_look_at_origin(cam_pos, 40.0, W, H, SPP, DEPTH)
    ref = None
    if refgpu.available():
        ref = refgpu.RefScene(tris, mats, cam_pos, sun, W, H, SPP, DEPTH)
        res["ref_scene"] = dict(nodes=ref.num_bvh_nodes, normal_mismatches=ref.normal_mismatches, build_ms=ref.build_ms)
        ours_cam = np.array(cam.origin.tuple() + cam.lower_left_corner.tuple() + cam.horizontal.tuple() + cam.vertical.tuple(), np.float32)
        res["ref_scene"]["camera_bits_equal"] = bool(np.array_equal(ours_cam.view(np.uint32), ref.camera12.view(np.uint32)))
        print("ref scene", res["ref_scene"], flush=True)

    def stage(name, fn):
        if name not in stages:
            return
        t0 = time.time()
        try:
            res[name] = fn()
        except Exception as e:  # noqa: BLE001
            res[name] = {"error": repr(e), "trace": traceback.format_exc()}
        res[name + "_s"] = round(time.time() - t0, 2)
        print(name, json.dumps(res[name])[:1500], flush=True)
        Path(args.out).write_text(json.dumps(res, indent=1))

    def s_bvh():
        nodes, recs = ctx.download_bvh()
        return check_bvh(nodes, recs, tris)

    def s_first():
        out = {}
        for jx, jy in ((0.5, 0.5), (0.25, 0.75)):
            ours = ctx.first_hits(cam, W, H, jx, jy)
            if ref is None:
                out[f"{jx},{jy}"] = dict(our_hits=int((ours[0] >= 0).sum()))
            else:
                out[f"{jx},{jy}"] = compare_hits(ours, ref.first_hits(jx, jy))
        return out

    def s_rays():
        lo = np.array(st.scene_lo[:]); hi = np.array(st.scene_hi[:])
        rays = random_rays(200000, 7, lo, hi)
        ours = ctx.trace_closest(rays)
        occ = ctx.trace_occluded(rays)
        out = dict(occluded=int(occ.sum()), closest_hits=int((ours[0] >= 0).sum()),
                   occluded_equals_closest=bool(np.array_equal(occ.astype(bool), ours[0] >= 0)))
        if ref is not None:
            out["vs_ref"] = compare_hits(ours, ref.trace(rays))
        return out

    def render(rng_mode, spp=SPP, seed=d.DEFAULT_SEED):
        params = d.make_params(W, H, spp, DEPTH, rng_mode)
        fs = ctx.render(cam, sun, params, seed=seed)
        img = ctx.readback_rgb8(spp, 2.0)
        acc = ctx.readback_accum()
        return img, acc, fs

    def s_lcg():
        img, acc, fs = render(d.RNG_REF_LCG)
        out = dict(device_ms=fs.device_ms, rays=int(fs.rays), primary=int(fs.primary_rays), bounce=int(fs.bounce_rays),
                   shadow=int(fs.shadow_rays), launches=fs.kernel_launches, mean=float(img.mean()),
                   nonzero_pixels=int((img.sum(2) > 0).sum()))
        d.write_png(OUT / "probe_lcg.png", img, W, H)
        if ref is not None:
            rimg, rms = ref.render()
            d.write_png(OUT / "probe_ref.png", rimg, W, H)
            same = (img == rimg).all(2)
            diff = np.abs(img.astype(np.int32) - rimg.astype(np.int32))
            out.update(ref_kernel_ms=rms, ref_mean=float(rimg.mean()), pixels=int(same.size), identical_pixels=int(same.sum()),
                       max_abs_diff=int(diff.max()), mean_abs_diff=float(diff.mean()),
                       differing_examples=[[int(a), int(b)] for a, b in np.argwhere(~same)[:6]])
        return out

    def s_philox():
        spp = 256
        params = d.make_params(W, H, spp, DEPTH, d.RNG_PHILOX)
        fs = ctx.render(cam, sun, params)
        img = ctx.readback_rgb8(spp, 2.0)
        acc = ctx.readback_accum() / spp
        out = dict(device_ms=fs.device_ms, rays=int(fs.rays), mrays_s=fs.rays / fs.device_ms / 1e3, mean=float(img.mean()))
        d.write_png(OUT / "probe_philox.png", img, W, H)
        # sample-sharded accumulate == single pass (same global sample indices)
        fs1 = ctx.render(cam, sun, params, sample_begin=0, sample_end=100)
        fs2 = ctx.render(cam, sun, params, sample_begin=100, sample_end=spp, keep_accum=True)
        acc2 = ctx.readback_accum() / spp
        out["shard_max_abs_diff"] = float(np.abs(acc2 - acc).max())
        # LCG replay at the same spp for the statistical comparison, and the reference itself
        if ref is not None:
            ref2 = refgpu.RefScene(tris, mats, cam_pos, sun, W, H, spp, DEPTH)
            rimg, rms = ref2.render()
            ref2.set_seed(1338)
            rimg2, _ = ref2.render()
            ref2.close()
            lin = lambda im: (im.astype(np.float64) / 255.0) ** 2.0   # undo gamma 2
            rmse = lambda a, b: float(np.sqrt(np.mean((lin(a) - lin(b)) ** 2)))
            out.update(ref_kernel_ms=rms, rmse_vs_ref=rmse(img, rimg), ref_seed_noise_rmse=rmse(rimg, rimg2),
                       mean_rel_diff=float(abs(lin(img).mean() - lin(rimg).mean()) / max(lin(rimg).mean(), 1e-12)),
                       mean_abs_diff_rgb8=float(np.abs(img.astype(np.int32) - rimg.astype(np.int32)).mean()),
                       ref_seed_mean_abs_diff_rgb8=float(np.abs(rimg2.astype(np.int32) - rimg.astype(np.int32)).mean()))
        return out

    def s_perf():
        out = {}
        tr, mt = d.synth_mesh(args.perf_tris)
        c2 = d.Context(0)
        s2 = c2.upload_scene(tr, mt)
        out["scene"] = dict(tris=s2.num_triangles, nodes=s2.num_wide_nodes, levels=s2.num_levels, build_ms=s2.build_ms,
                            upload_ms=s2.upload_ms, sah=s2.sah_cost)
        w, h, spp = 1920, 1080, 16
        cm = d.camera_look_at_origin(cam_pos, 40.0, w, h, spp, DEPTH)
        params = d.make_params(w, h, spp, DEPTH, d.RNG_PHILOX)
        c2.set_stage_timing(True)
        for it in range(3):
            fs = c2.render(cm, sun, params)
            out[f"run{it}"] = dict(device_ms=fs.device_ms, rays=int(fs.rays), primary=int(fs.primary_rays), bounce=int(fs.bounce_rays),
                                   shadow=int(fs.shadow_rays), mrays_s=fs.rays / fs.device_ms / 1e3, extend_ms=fs.extend_ms,
                                   primary_ms=fs.primary_ms, shade_ms=fs.shade_ms, launches=fs.kernel_launches, batches=fs.batches)
        img = c2.readback_rgb8(spp, 2.0)
        d.write_png(OUT / "probe_1m.png", img, w, h)
        c2.close()
        return out

    stage("bvh", s_bvh)
    stage("first", s_first)
    stage("rays", s_rays)
    stage("lcg", s_lcg)
    stage("philox", s_philox)
    stage("perf", s_perf)
    Path(args.out).write_text(json.dumps(res, indent=1))
    print("done", flush=True)


if __name__ == "__main__":
    main()
