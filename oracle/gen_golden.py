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

  --obj   (runs anywhere, no GPU): writes the OBJ/MTL/texture fixtures under tests/golden/obj/
          and flattens each with the reference's own triangle_mesh + build_gpu_scene compiled
          for the CPU (oracle/_ref/libref_host.so, oracle/ref_host_flatten.cpp)
          -> tests/golden/obj/<name>.ref.npz (triangles, materials, texture headers, pool)

Usage: python oracle/gen_golden.py --host | --gpu | --obj
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


OBJ_DIR = ROOT / "tests" / "golden" / "obj"


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

    # ---- spheres (with and without emissive ones) and the OBJ fixture: tests/test_gpu_scene_features.py
    import test_gpu_scene_features as tf
    fw, fh, fspp, fdepth = tf.W, tf.H, tf.SPP, tf.DEPTH
    out = {}
    for lights in (1, 0):
        tris, mats, sph = tf.sphere_scene(d, bool(lights))
        ref = refgpu.RefScene(tris, mats, np.array((350, 260, 1500), np.float32), tf.SUN, fw, fh, fspp, fdepth, spheres=sph)
        tri, t, u, v = ref.first_hits(0.5, 0.5)
        for k, a in zip(("tri", "t", "u", "v"), (tri, t, u, v)):
            out[f"sph{lights}_first_{k}"] = a
        out[f"sph{lights}_lcg_rgb"] = ref.render()[0]
        ref.close()
    ref = refgpu.RefScene(None, None, np.array((3, 4, 22), np.float32), np.array([-0.3, -0.5, -0.8], np.float32), fw, fh, fspp, fdepth,
                          obj=OBJ_DIR / "station.obj", obj_scale=2.5)
    out["station_lcg_rgb"] = ref.render()[0]
    ref.close()
    dst = ROOT / "gpurun_out" / "ref_gpu_features.npz"
    np.savez_compressed(dst, W=fw, H=fh, SPP=fspp, DEPTH=fdepth, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


OBJ_DIR = ROOT / "tests" / "golden" / "obj"


def ref_flatten(obj_path, scale=1.0):
    """The reference's OBJ -> GPUScene arrays (CPU build of its own code).  Call in a fresh
    process per scene: stb_image's vertical-flip flag is process-global in the reference."""
    L = C.CDLL(str(ROOT / "oracle" / "_ref" / "libref_host.so"))
    L.refflat_load_obj.restype = C.c_void_p
    L.refflat_load_obj.argtypes = [C.c_char_p, C.c_double]
    L.refflat_counts.argtypes = [C.c_void_p, C.c_void_p]
    L.refflat_copy.argtypes = [C.c_void_p] * 5
    L.refflat_free.argtypes = [C.c_void_p]
    h = L.refflat_load_obj(str(obj_path).encode(), scale)
    n = np.zeros(6, np.int32)
    L.refflat_counts(h, n.ctypes.data)
    tris = np.zeros((n[0], 29), np.float32)
    mats = np.zeros((n[1], 12), np.float32)
    th = np.zeros((n[3], 3), np.int32)
    pool = np.zeros(n[4], np.float32)
    L.refflat_copy(h, tris.ctypes.data, mats.ctypes.data, th.ctypes.data, pool.ctypes.data)
    L.refflat_free(h)
    return {"counts": n, "tris": tris.view(np.uint32), "mats": mats.view(np.uint32), "textures": th, "pool": pool.view(np.uint32)}


def write_obj_fixtures():
    """Small hand-written scenes that exercise every branch of the reference's loader."""
    import dsrt_b200 as d
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(7)
    # textures: a 5x3 PNG and a 4x4 PPM with distinct rows so a vertical flip is visible
    png = (rng.integers(0, 256, (3, 5, 3))).astype(np.uint8)
    d.write_png(OBJ_DIR / "panel.png", png, 5, 3)
    ppm = (rng.integers(0, 256, (4, 4, 3))).astype(np.uint8)
    d.write_ppm(OBJ_DIR / "hull.ppm", ppm, 4, 4)
    (OBJ_DIR / "station.mtl").write_text("""# materials for station.obj
newmtl hull_tex
Kd 0.5 0.4 0.3
map_Kd hull.ppm

newmtl panel_tex
Kd 1 1 1
Ks 0.9 0.9 0.9
map_Kd panel.png
newmtl glass
Kd 0.1 0.1 0.1
d 0.4
Ni 1.33
newmtl glass_bad_ior
d 0.5
Ni 25
newmtl gold
Kd 0.2 0.2 0.2
Ks 0.8 0.6 0.2
Ns 300
newmtl dull_metal
Ks 0.03 0.03 0.03
Kd 0.25 0.5 0.75
newmtl lamp
Ke 4 3.5 3
Kd 0 0 0
newmtl paint
Kd 0.123456789 0.987654321 1e-1
newmtl missing_tex
map_Kd not_there.png
newmtl unused
Kd 0 1 0
""")
    (OBJ_DIR / "station.obj").write_text("""# fixture: every index form, fans, all material classes
mtllib station.mtl
v -1.5 0 0.25
v 1.5 0 0.25
v 1.5 2 -0.125
v -1.5 2.0e0 -0.125
v 0 3.5 .5
v 2.75 1 1
v -2.75 1 1
v 0.1 0.2 5

vt 0 0
vt 1 0
vt 1 1
vt 0 1
vt 0.5 1.75
vn 0 0 1

f 1 2 3
usemtl hull_tex
f 1/1 2/2 3/3 4/4
usemtl panel_tex
f 4/4/1 3/3/1 5/5/1
usemtl glass
f 2//1 6//1 3//1
usemtl glass_bad_ior
f 1 4 7
usemtl gold
f 1/1/1 2/2/1 6/3/1 3/4/1 5/5/1
usemtl nowhere_defined
f 7 1 8
usemtl dull_metal
f 8 2 6
usemtl lamp
f 5 4 7
usemtl paint
f 6 5 8
f 3/9 6/1 8/2
usemtl missing_tex
f 1/1 8/2 2/3
usemtl hull_tex
f 7/1 8/2 5/3
f 1 2
f 0 1 2
usemtl gold
f 2 3 8
""")
    # edge: the only textured material is emissive -> no image_texture is constructed by the
    # reference, stb's flip flag stays off and the pool keeps the file's row order
    (OBJ_DIR / "lamp_only.mtl").write_text("""newmtl lit_panel
Ke 1 1 1
map_Kd panel.png
newmtl plain
Kd 0.3 0.3 0.3
""")
    (OBJ_DIR / "lamp_only.obj").write_text("""mtllib lamp_only.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
usemtl plain
f 1 2 3
usemtl lit_panel
f 1/1 3/3 4/4
""")
    return [("station", 2.5), ("lamp_only", 1.0)]


def obj_one(name, scale):
    r = ref_flatten(OBJ_DIR / f"{name}.obj", scale)
    np.savez_compressed(OBJ_DIR / f"{name}.ref.npz", scale=scale, **r)
    print("wrote", OBJ_DIR / f"{name}.ref.npz", r["counts"])


def obj():
    import subprocess
    for name, scale in write_obj_fixtures():
        subprocess.check_call([sys.executable, __file__, "--obj_one", name, str(scale)])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--obj", action="store_true")
    ap.add_argument("--obj_one", nargs=2)
    a = ap.parse_args()
    if a.host:
        host()
    if a.gpu:
        gpu()
    if a.obj:
        obj()
    if a.obj_one:
        obj_one(a.obj_one[0], float(a.obj_one[1]))
