"""Generates the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

The reference ships no tests or golden vectors, so the fixtures are
outputs of its own code, compiled unmodified into oracle/_ref/libref_gpu.so (oracle/Makefile):

  --host  (runs anywhere, no GPU): the reference's read_pose_file / rotate_yaw_deg_d /
          dnormalize / point_camera_at / camera::initialize for all 99 shipped poses
          -> tests/golden/pose_camera_golden.json (float bit patterns as hex)
  --gpu   (needs a GPU; run through gpurun): the reference's build_gpu_scene + render_kernel +
          scene_hit on the synthetic 10k-triangle station
          -> tests/golden/ref_gpu_10k.npz (first hits, random-ray hits, LCG image)
          (written to gpurun_out/ on the box and copied into tests/golden/ afterwards)

Usage: python oracle/gen_golden.py --host | --gpu
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

POSES = ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt"   # copy of the reference's orbit_sim file


def hexf(a):
    return [f"{x:08x}" for x in np.asarray(a, np.float32).view(np.uint32).ravel()]


def host():
    L = C.CDLL(str(ROOT / "oracle" / "_ref" / "libref_gpu.so"))
    L.refhost_read_pose_file.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    L.refhost_pose_to_model_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    L.refhost_camera.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    poses = np.zeros((256, 9))
    n = L.refhost_read_pose_file(str(POSES).encode(), poses.ctypes.data, 256)
    poses = poses[:n]
    out = {"source": "reference functions via oracle/ref_host_probe.cpp", "num_poses": n,
           "poses_sha256": hashlib.sha256(poses.tobytes()).hexdigest(), "frames": []}
    for i in range(n):
        cam, sun = np.zeros(3, np.float32), np.zeros(3, np.float32)
        sep = C.c_double()
        L.refhost_pose_to_model_frame(poses[i].ctypes.data, cam.ctypes.data, sun.ctypes.data, C.byref(sep))
        rec = {"frame": i, "sep_m": sep.value.hex(), "cam_in_model": hexf(cam), "sun_dir_model": hexf(sun), "cameras": {}}
        for (W, H) in ((800, 450), (1920, 1080), (400, 225)):
            cam12, uvw = np.zeros(12, np.float32), np.zeros(9, np.float32)
            L.refhost_camera(cam.ctypes.data, 40.0, W, H, cam12.ctypes.data, uvw.ctypes.data)
            rec["cameras"][f"{W}x{H}"] = {"cam12": hexf(cam12), "uvw": hexf(uvw)}
        out["frames"].append(rec)
    p = ROOT / "tests" / "golden" / "pose_camera_golden.json"
    p.write_text(json.dumps(out, indent=0))
    print("wrote", p, n, "poses")


def gpu():
    import dsrt_b200 as d
    import refgpu
    from gpu_probe import random_rays

    W, H, SPP, DEPTH = 200, 113, 8, 10
    poses = d.read_pose_file(POSES)
    out = {}
    tris, mats = d.synth_mesh(10000)
    out["mesh_sha256"] = np.frombuffer(hashlib.sha256(tris.tobytes()).digest(), np.uint8)
    for frame in (0, 98):
        cam, sun, _ = d.pose_to_model_frame(poses[frame])
        ref = refgpu.RefScene(tris, mats, cam, sun, W, H, SPP, DEPTH)
        assert ref.normal_mismatches == 0
        tri, t, u, v = ref.first_hits(0.5, 0.5)
        out[f"f{frame}_first_tri"] = tri
        out[f"f{frame}_first_t"] = t
        out[f"f{frame}_first_u"] = u
        out[f"f{frame}_first_v"] = v
        img, _ = ref.render()
        out[f"f{frame}_lcg_rgb"] = img
        if frame == 0:
            lo = np.stack([np.stack([tris[n][k] for k in "xyz"], 1) for n in ("v0", "v1", "v2")]).min((0, 1))
            hi = np.stack([np.stack([tris[n][k] for k in "xyz"], 1) for n in ("v0", "v1", "v2")]).max((0, 1))
            rays = random_rays(20000, 11, lo, hi)
            rt = ref.trace(rays)
            out["rays"] = rays
            out["rays_tri"], out["rays_t"], out["rays_u"], out["rays_v"] = rt
        ref.close()
    dst = ROOT / "gpurun_out" / "ref_gpu_10k.npz"
    dst.parent.mkdir(exist_ok=True)
    np.savez_compressed(dst, W=W, H=H, SPP=SPP, DEPTH=DEPTH, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    a = ap.parse_args()
    if a.host:
        host()
    if a.gpu:
        gpu()
