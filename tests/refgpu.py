"""ctypes access to oracle/_ref/libref_gpu.so -- the UNMODIFIED reference renderer compiled for
sm_100a (oracle/Makefile, oracle/ref_gpu_harness.cu).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_ref" / "libref_gpu.so"


def available() -> bool:
    return LIB.exists()


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(LIB))
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        _lib.ref_scene_create.restype = vp
        _lib.ref_scene_create.argtypes = [vp, i, vp, i, vp, f, i, i, i, i, vp]
        _lib.ref_scene_create_ex.restype = vp
        _lib.ref_scene_create_ex.argtypes = [vp, i, vp, i, vp, i, vp, f, i, i, i, i, vp]
        _lib.ref_scene_create_obj.restype = vp
        _lib.ref_scene_create_obj.argtypes = [C.c_char_p, C.c_double, vp, f, i, i, i, i, vp]
        _lib.ref_scene_arrays.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        _lib.ref_scene_info.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(f), vp]
        _lib.ref_scene_set_seed.argtypes = [vp, C.c_ulonglong]
        _lib.ref_render.argtypes = [vp, vp, C.POINTER(f)]
        _lib.ref_first_hits.argtypes = [vp, f, f, vp, vp, vp, vp]
        _lib.ref_trace.argtypes = [vp, vp, i, vp, vp, vp, vp]
        _lib.ref_scene_legacy.argtypes = [vp]
        _lib.ref_scene_legacy.restype = vp
        _lib.ref_scene_destroy.argtypes = [vp]
        _lib.ref_scene_destroy.restype = None
    return _lib


class RefScene:
    """The reference's own build_gpu_scene + render_kernel on the same triangles."""

    def __init__(self, tris, mats, lookfrom, sun_dir, width, height, spp, max_depth, vfov=40.0, spheres=None, obj=None, obj_scale=1.0):
        """Triangles/materials (+ optional spheres, 24-byte GPUSphere records) become reference
        scene-graph objects; or, with obj=<path>, the reference's own triangle_mesh loads the OBJ."""
        L = lib()
        lf = np.ascontiguousarray(lookfrom, dtype=np.float32)
        sd = np.ascontiguousarray(sun_dir, dtype=np.float32)
        self.W, self.H = width, height
        if obj is not None:
            self.h = L.ref_scene_create_obj(str(obj).encode(), float(obj_scale), lf.ctypes.data, vfov,
                                            width, height, spp, max_depth, sd.ctypes.data)
        else:
            t = np.ascontiguousarray(tris)
            m = np.ascontiguousarray(mats)
            sp = np.ascontiguousarray(spheres) if spheres is not None and len(spheres) else None
            self.h = L.ref_scene_create_ex(t.ctypes.data if len(t) else None, len(t), m.ctypes.data, len(m),
                                           sp.ctypes.data if sp is not None else None, len(sp) if sp is not None else 0,
                                           lf.ctypes.data, vfov, width, height, spp, max_depth, sd.ctypes.data)
        if not self.h:
            raise RuntimeError("ref_scene_create failed")
        nb, nm, ms = C.c_int(), C.c_int(), C.c_float()
        cam = np.zeros(12, np.float32)
        L.ref_scene_info(self.h, C.byref(nb), C.byref(nm), C.byref(ms), cam.ctypes.data)
        self.num_bvh_nodes, self.normal_mismatches, self.build_ms = nb.value, nm.value, ms.value
        self.camera12 = cam

    def arrays(self):
        """The flattened arrays build_gpu_scene uploaded: dict of uint32/int32 word views."""
        n = np.zeros(6, np.int32)
        assert lib().ref_scene_arrays(self.h, n.ctypes.data, None, None, None, None, None) == 0
        tris = np.zeros((n[0], 29), np.uint32)
        mats = np.zeros((n[1], 12), np.uint32)
        sph = np.zeros((n[2], 6), np.uint32)
        th = np.zeros((n[3], 3), np.int32)
        pool = np.zeros(n[4], np.uint32)
        assert lib().ref_scene_arrays(self.h, n.ctypes.data, tris.ctypes.data, mats.ctypes.data, sph.ctypes.data,
                                      th.ctypes.data, pool.ctypes.data) == 0
        return {"counts": n, "tris": tris, "mats": mats, "spheres": sph, "textures": th, "pool": pool}

    def legacy_scene_ptr(self) -> int:
        """Address of the reference's GPUScene (host struct holding device pointers)."""
        return lib().ref_scene_legacy(self.h)

    def set_seed(self, seed: int):
        assert lib().ref_scene_set_seed(self.h, seed) == 0

    def render(self):
        rgb = np.empty((self.H, self.W, 3), np.uint8)
        ms = C.c_float()
        rc = lib().ref_render(self.h, rgb.ctypes.data, C.byref(ms))
        if rc != 0:
            raise RuntimeError(f"ref_render failed {rc}")
        return rgb, ms.value

    def first_hits(self, jx=0.5, jy=0.5):
        n = self.W * self.H
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        rc = lib().ref_first_hits(self.h, jx, jy, tri.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ref_first_hits failed {rc}")
        return tri, t, u, v

    def trace(self, rays):
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = len(r)
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        rc = lib().ref_trace(self.h, r.ctypes.data, n, tri.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ref_trace failed {rc}")
        return tri, t, u, v

    def close(self):
        if self.h:
            lib().ref_scene_destroy(self.h)
            self.h = None

    __del__ = close
