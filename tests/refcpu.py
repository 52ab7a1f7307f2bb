"""ctypes access to oracle/_ref/libref_cpu.so -- the reference's CPU classes (camera, bvh_node,
triangle, aabb ...) driven by oracle/ref_cpu_baseline.cpp.  Bench/test infrastructure only."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_ref" / "libref_cpu.so"


def available() -> bool:
    return LIB.exists()


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(LIB))
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.refcpu_create.restype = vp
        L.refcpu_create.argtypes = [vp, i, vp, i, C.POINTER(C.c_double)]
        L.refcpu_destroy.argtypes = [vp]
        L.refcpu_destroy.restype = None
        L.refcpu_render.restype = C.c_double
        L.refcpu_render.argtypes = [vp, vp, f, i, i, i, i, vp, vp, C.c_ulonglong, i, i, i, i, i, vp, vp]
        _lib = L
    return _lib


class RefCpuScene:
    def __init__(self, tris, mats):
        t, m = np.ascontiguousarray(tris), np.ascontiguousarray(mats)
        b = C.c_double()
        self.h = lib().refcpu_create(t.ctypes.data, len(t), m.ctypes.data, len(m), C.byref(b))
        self.build_s = b.value

    def render(self, lookfrom, sun_dir, W, H, spp, max_depth, step_x=1, step_y=1, s_begin=0, s_end=None,
               threads=0, sun_rad=(100000.0, 95000.0, 90000.0), seed=1337, accum=None, vfov=40.0):
        """Returns (seconds, counts[primary, bounce, shadow])."""
        lf = np.ascontiguousarray(lookfrom, np.float32)
        sd = np.ascontiguousarray(sun_dir, np.float32)
        sr = np.ascontiguousarray(sun_rad, np.float32)
        counts = np.zeros(3, np.uint64)
        if s_end is None:
            s_end = spp
        ap = accum.ctypes.data if accum is not None else None
        secs = lib().refcpu_render(self.h, lf.ctypes.data, vfov, W, H, spp, max_depth, sd.ctypes.data, sr.ctypes.data,
                                   seed, step_x, step_y, s_begin, s_end, threads, ap, counts.ctypes.data)
        return secs, counts

    def close(self):
        if getattr(self, "h", None):
            lib().refcpu_destroy(self.h)
            self.h = None

    __del__ = close
