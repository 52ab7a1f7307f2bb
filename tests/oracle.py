"""ctypes access to oracle/liboracle.so -- the CPU restatement of the reference's algorithm.
Test infrastructure only (see oracle/oracle.h)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "liboracle.so"

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "liboracle.so"])
        L = C.CDLL(str(LIB))
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.or_scene_create.restype = vp
        L.or_scene_create.argtypes = [vp, i, vp, i]
        L.or_scene_destroy.argtypes = [vp]
        L.or_scene_destroy.restype = None
        L.or_scene_set_spheres.argtypes = [vp, vp, i]
        L.or_scene_set_spheres.restype = None
        L.or_scene_set_textures.argtypes = [vp, vp, i, vp, i]
        L.or_scene_set_textures.restype = None
        L.or_scene_num_nodes.argtypes = [vp]
        L.or_scene_max_leaf.argtypes = [vp]
        L.or_first_hits.argtypes = [vp, vp, i, i, f, f, vp, vp, vp, vp]
        L.or_first_hits.restype = None
        L.or_primary_hits.argtypes = [vp, vp, i, i, i, i, C.c_uint64, i, vp, vp, vp, vp]
        L.or_primary_hits.restype = None
        L.or_trace.argtypes = [vp, vp, i, vp, vp, vp, vp]
        L.or_trace.restype = None
        L.or_occluded.argtypes = [vp, vp, i, vp]
        L.or_occluded.restype = None
        L.or_render.argtypes = [vp, vp, i, i, i, i, vp, vp, i, C.c_uint64, i, f, i, i, vp, vp, vp, vp]
        L.or_render.restype = None
        L.or_tonemap.argtypes = [vp, i, i, i, f, vp]
        L.or_tonemap.restype = None
        L.or_read_pose_file.argtypes = [C.c_char_p, vp, i]
        L.or_pose_to_model_frame.argtypes = [vp, vp, vp, C.POINTER(C.c_double)]
        L.or_pose_to_model_frame.restype = None
        L.or_camera_init.argtypes = [vp, f, i, i, vp, vp]
        L.or_camera_init.restype = None
        L.or_set_threads.argtypes = [i]
        L.or_set_threads.restype = None
        L.or_philox4x32_10.argtypes = [vp, vp, vp]
        L.or_philox4x32_10.restype = None
        _lib = L
    return _lib


def read_pose_file(path):
    n = lib().or_read_pose_file(str(path).encode(), None, 0)
    if n < 0:
        raise OSError(path)
    out = np.zeros((n, 9))
    if n:
        lib().or_read_pose_file(str(path).encode(), out.ctypes.data, n)
    return out


def pose_to_model_frame(pose9):
    p = np.ascontiguousarray(pose9, dtype=np.float64)
    cam, sun = np.zeros(3, np.float32), np.zeros(3, np.float32)
    sep = C.c_double()
    lib().or_pose_to_model_frame(p.ctypes.data, cam.ctypes.data, sun.ctypes.data, C.byref(sep))
    return cam, sun, sep.value


def camera_init(lookfrom, vfov, W, H):
    lf = np.ascontiguousarray(lookfrom, dtype=np.float32)
    cam12, uvw = np.zeros(12, np.float32), np.zeros(9, np.float32)
    lib().or_camera_init(lf.ctypes.data, vfov, W, H, cam12.ctypes.data, uvw.ctypes.data)
    return cam12, uvw


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().or_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


class Scene:
    def __init__(self, tris, mats, spheres=None, textures=None, texture_pool=None):
        t = np.ascontiguousarray(tris)
        m = np.ascontiguousarray(mats)
        assert t.dtype.itemsize == 116 and m.dtype.itemsize == 48
        self._keep = (t, m)
        self.h = lib().or_scene_create(t.ctypes.data, len(t), m.ctypes.data, len(m))
        if spheres is not None and len(spheres):
            sp = np.ascontiguousarray(spheres)
            assert sp.dtype.itemsize == 24
            lib().or_scene_set_spheres(self.h, sp.ctypes.data, len(sp))
        if textures is not None and len(textures):
            th = np.ascontiguousarray(textures)
            tp = np.ascontiguousarray(texture_pool, np.float32)
            assert th.dtype.itemsize == 12
            lib().or_scene_set_textures(self.h, th.ctypes.data, len(th), tp.ctypes.data, len(tp))
        self.num_nodes = lib().or_scene_num_nodes(self.h)
        self.max_leaf = lib().or_scene_max_leaf(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib().or_scene_destroy(self.h)
            self.h = None

    __del__ = close

    def first_hits(self, cam12, W, H, jx=0.5, jy=0.5):
        n = W * H
        c = np.ascontiguousarray(cam12, np.float32)
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        lib().or_first_hits(self.h, c.ctypes.data, W, H, jx, jy, tri.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
        return tri, t, u, v

    def primary_hits(self, cam12, W, H, spp, sample=0, seed=1337, rng_mode=0):
        """First hit of sample `sample` of every pixel under the estimator's jitter (rng_mode 1: LCG, sample 0 only)."""
        assert rng_mode == 0 or sample == 0
        n = W * H
        c = np.ascontiguousarray(cam12, np.float32)
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        lib().or_primary_hits(self.h, c.ctypes.data, W, H, spp, sample, seed, rng_mode,
                              tri.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
        return tri, t, u, v

    def trace(self, rays):
        r = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = len(r)
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        lib().or_trace(self.h, r.ctypes.data, n, tri.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
        return tri, t, u, v

    def occluded(self, rays):
        r = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        occ = np.empty(len(r), np.uint8)
        lib().or_occluded(self.h, r.ctypes.data, len(r), occ.ctypes.data)
        return occ

    def render(self, cam12, W, H, spp, max_depth, sun_dir, sun_rad=(100000.0, 95000.0, 90000.0), sun_enabled=True,
               seed=1337, rng_mode=1, gamma=2.0, s_begin=0, s_end=None, accum=None, lcg_state=None):
        """Returns (accum float32 HxWx3 bottom-up SUM, rgb8 HxWx3 top-down, ray_counts[3])."""
        if s_end is None:
            s_end = spp
        c = np.ascontiguousarray(cam12, np.float32)
        sd = np.ascontiguousarray(sun_dir, np.float32)
        sr = np.ascontiguousarray(sun_rad, np.float32)
        if accum is None:
            accum = np.zeros((H, W, 3), np.float32)
        rgb = np.empty((H, W, 3), np.uint8)
        counts = np.zeros(3, np.uint64)
        ls = lcg_state.ctypes.data if lcg_state is not None else None
        lib().or_render(self.h, c.ctypes.data, W, H, spp, max_depth, sd.ctypes.data, sr.ctypes.data, int(sun_enabled),
                        seed, rng_mode, gamma, s_begin, s_end, accum.ctypes.data, rgb.ctypes.data, ls, counts.ctypes.data)
        return accum, rgb, counts


def tonemap(accum, spp, gamma=2.0):
    a = np.ascontiguousarray(accum, np.float32)
    H, W, _ = a.shape
    rgb = np.empty((H, W, 3), np.uint8)
    lib().or_tonemap(a.ctypes.data, W, H, spp, gamma, rgb.ctypes.data)
    return rgb
