"""The synthetic station and the GPUScene record layouts WITHOUT the product library:
oracle/libsynthmesh.so is the generator's source compiled on its own (oracle/Makefile), so that the
baseline legs of bench.py (`--impl reference`, cpu_baseline) never map libdsrt.so.  Bench/test
infrastructure only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "libsynthmesh.so"

vec3 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4")])
triangle_dtype = np.dtype(
    [("v0", vec3), ("v1", vec3), ("v2", vec3), ("n0", vec3), ("n1", vec3), ("n2", vec3),
     ("uv0", vec3), ("uv1", vec3), ("uv2", vec3), ("material_id", "<i4"), ("albedo_tex", "<i4")]
)
material_dtype = np.dtype(
    [("type", "<i4"), ("albedo_tex", "<i4"), ("pad0", "<i4"), ("pad1", "<i4"),
     ("albedo", vec3), ("emissive", vec3), ("fuzz", "<f4"), ("ref_idx", "<f4")]
)
assert triangle_dtype.itemsize == 116 and material_dtype.itemsize == 48

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "libsynthmesh.so"])
        _lib = C.CDLL(str(LIB))
        _lib.dsrt_synth_mesh.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    return _lib


def synth_mesh(ntris: int, seed: int = 1337, diffuse_only: bool = False):
    tris = np.zeros(ntris, dtype=triangle_dtype)
    mats = np.zeros(32, dtype=material_dtype)
    nm = lib().dsrt_synth_mesh(ntris, seed, int(diffuse_only), tris.ctypes.data, mats.ctypes.data, len(mats))
    if nm < 0:
        raise RuntimeError("dsrt_synth_mesh failed")
    return tris, mats[:nm].copy()
