"""ctypes binding of libdsrt.so -- the C ABI of the B200 render path (include/dsrt.h).

This package is plumbing only: it loads the in-tree shared library built from `csrc/` and
`host/` and mirrors the C structs.  It contains no rendering code and no fallback: if the
library is missing, importing raises; if no CUDA device is usable, `Context()` raises.

The reference interface this stands in for is the host->CUDA boundary of
Ofredy/Deep-Space-Ray-Tracer: build_gpu_scene / gpu_render_scene / free_gpu_scene
(reference inc/gpu_scene_builder.h:72-73, src/gpu_render.cu:1037-1108).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
# DSRT_LIB selects another build of the same library (kernel-variant experiments only).
LIB_PATH = Path(os.environ["DSRT_LIB"]) if os.environ.get("DSRT_LIB") else _HERE / "libdsrt.so"

if not LIB_PATH.exists():
    raise ImportError(
        f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C deep-space-ray-tracer_b200`). There is no CPU fallback."
    )

lib = C.CDLL(str(LIB_PATH))

# ---- error codes ---------------------------------------------------------------------------
OK, EINVAL, ECUDA, ENOSCENE, EUNSUPPORTED, EIO, ENOMEM = 0, -1, -2, -3, -4, -5, -6
RNG_PHILOX, RNG_REF_LCG = 0, 1
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT = 0, 1, 2, 3

# ---- numpy mirrors of the POD records (byte-compatible with include/dsrt_types.h) ------------
vec3 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4")])
triangle_dtype = np.dtype(
    [("v0", vec3), ("v1", vec3), ("v2", vec3), ("n0", vec3), ("n1", vec3), ("n2", vec3),
     ("uv0", vec3), ("uv1", vec3), ("uv2", vec3), ("material_id", "<i4"), ("albedo_tex", "<i4")]
)
material_dtype = np.dtype(
    [("type", "<i4"), ("albedo_tex", "<i4"), ("pad0", "<i4"), ("pad1", "<i4"),
     ("albedo", vec3), ("emissive", vec3), ("fuzz", "<f4"), ("ref_idx", "<f4")]
)
sphere_dtype = np.dtype([("center", vec3), ("radius", "<f4"), ("material_id", "<i4"), ("pad", "<i4")])
texture_header_dtype = np.dtype([("width", "<i4"), ("height", "<i4"), ("offset", "<i4")])
assert triangle_dtype.itemsize == 116 and material_dtype.itemsize == 48
assert sphere_dtype.itemsize == 24 and texture_header_dtype.itemsize == 12


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    def tuple(self):
        return (self.x, self.y, self.z)


class Camera(C.Structure):
    _fields_ = [("origin", Vec3), ("lower_left_corner", Vec3), ("horizontal", Vec3), ("vertical", Vec3),
                ("u", Vec3), ("v", Vec3), ("w", Vec3), ("lens_radius", C.c_float),
                ("image_width", C.c_int32), ("image_height", C.c_int32),
                ("samples_per_pixel", C.c_int32), ("max_depth", C.c_int32)]


class RenderParams(C.Structure):
    _fields_ = [("img_width", C.c_int32), ("img_height", C.c_int32), ("samples_per_pixel", C.c_int32),
                ("max_depth", C.c_int32), ("use_bvh", C.c_int32), ("rng_mode", C.c_int32),
                ("tile_size", C.c_int32), ("pad0", C.c_int32), ("gamma", C.c_float),
                ("exposure", C.c_float), ("env_rotation", C.c_float), ("pad1", C.c_float)]


class SceneStats(C.Structure):
    _fields_ = [("num_triangles", C.c_int32), ("num_wide_nodes", C.c_int32), ("num_levels", C.c_int32),
                ("max_leaf_tris", C.c_int32), ("build_ms", C.c_float), ("upload_ms", C.c_float),
                ("sah_cost", C.c_float), ("scene_lo", C.c_float * 3), ("scene_hi", C.c_float * 3),
                ("node_bytes", C.c_uint64), ("tri_bytes", C.c_uint64)]


class FrameStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("bounce_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("paths", C.c_uint64), ("kernel_launches", C.c_uint32), ("batches", C.c_uint32),
                ("device_ms", C.c_float), ("extend_ms", C.c_float), ("primary_ms", C.c_float),
                ("shade_ms", C.c_float), ("culled_primary_rays", C.c_uint64)]

    @property
    def rays(self):
        return self.primary_rays + self.bounce_rays + self.shadow_rays


assert C.sizeof(Camera) == 104 and C.sizeof(RenderParams) == 48

# Every symbol include/dsrt.h and include/dsrt_legacy.h declare (tests check they all resolve).
SYMBOLS = [
    "dsrt_create", "dsrt_destroy", "dsrt_last_error", "dsrt_stream",
    "dsrt_scene_upload", "dsrt_scene_upload_device", "dsrt_scene_stats", "dsrt_scene_download",
    "dsrt_render_frame", "dsrt_frame_stats", "dsrt_set_stage_timing", "dsrt_set_samples_per_batch",
    "dsrt_render_frame_async", "dsrt_readback_rgb8_async", "dsrt_frame_wait", "dsrt_primary_hits",
    "dsrt_set_bvh_builder",
    "dsrt_host_alloc", "dsrt_host_free", "dsrt_box_screen_bounds", "dsrt_readback_rgb8", "dsrt_readback_accum", "dsrt_accum_device", "dsrt_upload_accum", "dsrt_reduce_accum", "dsrt_clear_accum",
    "dsrt_trace_closest", "dsrt_trace_occluded", "dsrt_first_hits",
    "dsrt_trace_closest_device", "dsrt_trace_occluded_device",
    "dsrt_read_pose_file", "dsrt_pose_to_model_frame", "dsrt_camera_look_at_origin",
    "dsrt_synth_mesh", "dsrt_mesh_load_obj", "dsrt_mesh_view", "dsrt_mesh_free",
    "dsrt_write_ppm", "dsrt_write_png", "dsrt_version",
    "gpu_render_scene", "dsrt_scene_upload_legacy",
]

_vp, _i, _f = C.c_void_p, C.c_int, C.c_float
lib.dsrt_create.argtypes = [C.POINTER(_vp), _i]
lib.dsrt_destroy.argtypes = [_vp]
lib.dsrt_destroy.restype = None
lib.dsrt_last_error.argtypes = [_vp]
lib.dsrt_last_error.restype = C.c_char_p
lib.dsrt_stream.argtypes = [_vp]
lib.dsrt_stream.restype = _vp
lib.dsrt_scene_upload.argtypes = [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i]
lib.dsrt_scene_upload_device.argtypes = [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i]
lib.dsrt_scene_upload_legacy.argtypes = [_vp, _vp, _i, C.POINTER(_i)]
lib.dsrt_scene_stats.argtypes = [_vp, C.POINTER(SceneStats)]
lib.dsrt_scene_download.argtypes = [_vp, _vp, C.c_size_t, _vp, C.c_size_t]
lib.dsrt_render_frame.argtypes = [_vp, C.POINTER(Camera), C.POINTER(_f), C.POINTER(_f), _i,
                                  C.POINTER(RenderParams), C.c_uint64, _i, _i, C.POINTER(_i), _i]
lib.dsrt_frame_stats.argtypes = [_vp, C.POINTER(FrameStats)]
lib.dsrt_render_frame_async.argtypes = [_vp, _i, C.POINTER(Camera), C.POINTER(_f), C.POINTER(_f), _i,
                                        C.POINTER(RenderParams), C.c_uint64, _i, _i, C.POINTER(_i), _i]
lib.dsrt_readback_rgb8_async.argtypes = [_vp, _i, _i, _f, _vp, C.c_size_t]
lib.dsrt_frame_wait.argtypes = [_vp, _i, C.POINTER(FrameStats)]
lib.dsrt_primary_hits.argtypes = [_vp, C.POINTER(Camera), C.POINTER(RenderParams), C.c_uint64, _i, _vp, _vp, _vp, _vp, _vp, _vp]
lib.dsrt_set_stage_timing.argtypes = [_vp, _i]
lib.dsrt_set_samples_per_batch.argtypes = [_vp, _i]
lib.dsrt_set_bvh_builder.argtypes = [_vp, _i]
lib.dsrt_readback_rgb8.argtypes = [_vp, _i, _f, _vp, C.c_size_t]
lib.dsrt_readback_accum.argtypes = [_vp, _vp, C.c_size_t]
lib.dsrt_accum_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]
lib.dsrt_upload_accum.argtypes = [_vp, _vp, C.c_size_t]
lib.dsrt_clear_accum.argtypes = [_vp, _i, _i]
lib.dsrt_reduce_accum.argtypes = [_vp, _vp, _i]
lib.dsrt_trace_closest.argtypes = [_vp, _vp, _i, _vp, _vp, _vp, _vp]
lib.dsrt_trace_occluded.argtypes = [_vp, _vp, _i, _vp]
lib.dsrt_first_hits.argtypes = [_vp, C.POINTER(Camera), _i, _i, _f, _f, _vp, _vp, _vp, _vp]
lib.dsrt_trace_closest_device.argtypes = [_vp, _vp, _i, _vp]
lib.dsrt_trace_occluded_device.argtypes = [_vp, _vp, _i, _vp]
lib.dsrt_read_pose_file.argtypes = [C.c_char_p, _vp, _i]
lib.dsrt_pose_to_model_frame.argtypes = [_vp, _vp, _vp, C.POINTER(C.c_double)]
lib.dsrt_pose_to_model_frame.restype = None
lib.dsrt_camera_look_at_origin.argtypes = [_vp, _f, _i, _i, _i, _i, C.POINTER(Camera)]
lib.dsrt_camera_look_at_origin.restype = None
lib.dsrt_synth_mesh.argtypes = [_i, C.c_uint32, _i, _vp, _vp, _i]
lib.dsrt_mesh_load_obj.argtypes = [C.c_char_p, C.c_double, C.POINTER(_vp)]
lib.dsrt_mesh_view.argtypes = [_vp, _vp]
lib.dsrt_mesh_free.argtypes = [_vp]
lib.dsrt_mesh_free.restype = None
lib.dsrt_write_ppm.argtypes = [C.c_char_p, _vp, _i, _i]
lib.dsrt_write_png.argtypes = [C.c_char_p, _vp, _i, _i]
lib.dsrt_version.restype = C.c_char_p

SUN_RADIANCE = (100000.0, 95000.0, 90000.0)   # reference gpu_scene_builder.cpp:598
DEFAULT_SEED = 1337                            # reference gpu_scene_builder.cpp:583


class DsrtError(RuntimeError):
    pass


def version() -> str:
    return lib.dsrt_version().decode()


# ---- host side of the path (no GPU) -----------------------------------------------------------
def read_pose_file(path) -> np.ndarray:
    """reference read_pose_file (src/main.cpp:139-173): (n, 9) float64."""
    n = lib.dsrt_read_pose_file(os.fsencode(str(path)), None, 0)
    if n < 0:
        raise DsrtError(f"cannot read pose file {path}")
    out = np.zeros((n, 9), dtype=np.float64)
    if n:
        lib.dsrt_read_pose_file(os.fsencode(str(path)), out.ctypes.data, n)
    return out


def pose_to_model_frame(pose9):
    """reference main.cpp:334-357: returns (cam_in_model f32[3], sun_dir_model f32[3], sep_m)."""
    p = np.ascontiguousarray(pose9, dtype=np.float64)
    cam = np.zeros(3, np.float32)
    sun = np.zeros(3, np.float32)
    sep = C.c_double(0)
    lib.dsrt_pose_to_model_frame(p.ctypes.data, cam.ctypes.data, sun.ctypes.data, C.byref(sep))
    return cam, sun, sep.value


def camera_look_at_origin(lookfrom, vfov=40.0, width=800, height=450, spp=1000, max_depth=50) -> Camera:
    """reference point_camera_at + camera::initialize (main.cpp:178-187, camera.h:91-116)."""
    lf = np.ascontiguousarray(lookfrom, dtype=np.float32)
    cam = Camera()
    lib.dsrt_camera_look_at_origin(lf.ctypes.data, vfov, width, height, spp, max_depth, C.byref(cam))
    return cam


def synth_mesh(ntris: int, seed: int = 1337, diffuse_only: bool = False):
    """Deterministic synthetic station: (triangles, materials) structured arrays."""
    tris = np.zeros(ntris, dtype=triangle_dtype)
    mats = np.zeros(32, dtype=material_dtype)
    nm = lib.dsrt_synth_mesh(ntris, seed, int(diffuse_only), tris.ctypes.data, mats.ctypes.data, len(mats))
    if nm < 0:
        raise DsrtError("dsrt_synth_mesh failed")
    return tris, mats[:nm].copy()


class MeshView(C.Structure):
    _fields_ = [("triangles", _vp), ("num_triangles", C.c_int32), ("materials", _vp), ("num_materials", C.c_int32),
                ("textures", _vp), ("num_textures", C.c_int32), ("texture_pool", _vp), ("texture_pool_floats", C.c_int32),
                ("num_vertices", C.c_int32), ("num_uvs", C.c_int32), ("skipped_faces", C.c_int32),
                ("textures_failed", C.c_int32), ("textures_flipped", C.c_int32), ("textures_unsupported", C.c_int32)]


def load_obj(path, scale: float = 1.0) -> dict:
    """OBJ + MTL + map_Kd textures through dsrt_mesh_load_obj (host/obj_loader.cpp): dict with
    the structured arrays `tris`, `mats`, `textures`, the float `texture_pool` and the counters."""
    h = _vp()
    rc = lib.dsrt_mesh_load_obj(os.fsencode(str(path)), float(scale), C.byref(h))
    if rc != OK:
        raise DsrtError(f"cannot load OBJ {path} ({rc})")
    try:
        v = MeshView()
        lib.dsrt_mesh_view(h, C.byref(v))

        def grab(ptr, n, dtype):
            if not n:
                return np.zeros(0, dtype)
            buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=n).copy()

        return {"tris": grab(v.triangles, v.num_triangles, triangle_dtype),
                "mats": grab(v.materials, v.num_materials, material_dtype),
                "textures": grab(v.textures, v.num_textures, texture_header_dtype),
                "texture_pool": grab(v.texture_pool, v.texture_pool_floats, np.float32),
                "num_vertices": v.num_vertices, "num_uvs": v.num_uvs, "skipped_faces": v.skipped_faces,
                "textures_failed": v.textures_failed, "textures_flipped": bool(v.textures_flipped),
                "textures_unsupported": v.textures_unsupported}
    finally:
        lib.dsrt_mesh_free(h)


def make_params(width, height, spp, max_depth, rng_mode=RNG_PHILOX, gamma=2.0) -> RenderParams:
    p = RenderParams()
    p.img_width, p.img_height, p.samples_per_pixel, p.max_depth = width, height, spp, max_depth
    p.use_bvh, p.rng_mode, p.gamma, p.exposure = 1, rng_mode, gamma, 50.0
    return p


def write_ppm(path, rgb, width, height):
    a = np.ascontiguousarray(rgb, dtype=np.uint8)
    if lib.dsrt_write_ppm(os.fsencode(str(path)), a.ctypes.data, width, height) != OK:
        raise DsrtError(f"cannot write {path}")


def write_png(path, rgb, width, height):
    a = np.ascontiguousarray(rgb, dtype=np.uint8)
    if lib.dsrt_write_png(os.fsencode(str(path)), a.ctypes.data, width, height) != OK:
        raise DsrtError(f"cannot write {path}")


# ---- device side --------------------------------------------------------------------------------
class Context:
    """One GPU, one stream, one scene.  Raises DsrtError when no CUDA device is usable."""

    def __init__(self, device: int = 0):
        h = _vp()
        rc = lib.dsrt_create(C.byref(h), device)
        if rc != OK:
            raise DsrtError(f"dsrt_create({device}) failed ({rc}): {lib.dsrt_last_error(None).decode()}")
        self._h = h
        self.width = self.height = 0

    def close(self):
        if getattr(self, "_h", None):
            lib.dsrt_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != OK:
            raise DsrtError(f"dsrt error {rc}: {lib.dsrt_last_error(self._h).decode()}")

    @property
    def stream(self) -> int:
        return lib.dsrt_stream(self._h) or 0

    def upload_scene(self, tris: np.ndarray, mats: np.ndarray, spheres: np.ndarray | None = None,
                     textures: np.ndarray | None = None, texture_pool: np.ndarray | None = None) -> SceneStats:
        """tris/mats/spheres/textures: structured arrays in the reference's GPUScene layouts;
        texture_pool: float32 RGB pool the texture headers index (gpu_scene.h:28-32)."""
        t = np.ascontiguousarray(tris, dtype=triangle_dtype)
        m = np.ascontiguousarray(mats, dtype=material_dtype)
        sp = np.ascontiguousarray(spheres if spheres is not None else [], dtype=sphere_dtype)
        th = np.ascontiguousarray(textures if textures is not None else [], dtype=texture_header_dtype)
        tp = np.ascontiguousarray(texture_pool if texture_pool is not None else [], dtype=np.float32)
        self._ck(lib.dsrt_scene_upload(self._h, t.ctypes.data if len(t) else None, len(t), m.ctypes.data, len(m),
                                       sp.ctypes.data if len(sp) else None, len(sp),
                                       th.ctypes.data if len(th) else None, len(th),
                                       tp.ctypes.data if len(tp) else None, len(tp)))
        return self.scene_stats()

    def upload_scene_device(self, tris_ptr: int, ntris: int, mats_ptr: int, nmats: int,
                            spheres_ptr: int = 0, nspheres: int = 0) -> SceneStats:
        self._ck(lib.dsrt_scene_upload_device(self._h, tris_ptr, ntris, mats_ptr, nmats, spheres_ptr or None, nspheres,
                                              None, 0, None, 0))
        return self.scene_stats()

    def upload_scene_legacy(self, legacy_scene_ptr: int, reference_flat_boxes: bool = False):
        """From a reference GPUScene (host struct of device pointers, include/dsrt_legacy.h); returns
        (SceneStats, number of triangles masked because the reference's own BVH hides them)."""
        n = _i(0)
        self._ck(lib.dsrt_scene_upload_legacy(self._h, legacy_scene_ptr, int(reference_flat_boxes), C.byref(n)))
        return self.scene_stats(), n.value

    def scene_stats(self) -> SceneStats:
        st = SceneStats()
        self._ck(lib.dsrt_scene_stats(self._h, C.byref(st)))
        return st

    def download_bvh(self):
        st = self.scene_stats()
        nodes = np.zeros((st.num_wide_nodes, 24), dtype=np.uint32)      # 96-byte nodes (csrc/bvh8.cuh)
        recs = np.zeros((st.num_triangles, 12), dtype=np.float32)
        self._ck(lib.dsrt_scene_download(self._h, nodes.ctypes.data, nodes.nbytes, recs.ctypes.data, recs.nbytes))
        return nodes, recs

    def render(self, camera: Camera, sun_dir, params: RenderParams, seed=DEFAULT_SEED, sample_begin=0,
               sample_end=None, rect=None, keep_accum=False, sun_radiance=SUN_RADIANCE, sun_enabled=True) -> FrameStats:
        sd = (_f * 3)(*[float(x) for x in sun_dir])
        sr = (_f * 3)(*[float(x) for x in sun_radiance])
        if sample_end is None:
            sample_end = max(1, params.samples_per_pixel)
        r = (C.c_int * 4)(*rect) if rect is not None else None
        self._ck(lib.dsrt_render_frame(self._h, C.byref(camera), sd, sr, int(sun_enabled), C.byref(params),
                                       seed, sample_begin, sample_end, r, int(keep_accum)))
        self.width, self.height = params.img_width, params.img_height
        return self.frame_stats()

    def render_async(self, slot: int, camera: Camera, sun_dir, params: RenderParams, seed=DEFAULT_SEED, sample_begin=0,
                     sample_end=None, rect=None, keep_accum=False, sun_radiance=SUN_RADIANCE, sun_enabled=True):
        """Enqueue a frame on `slot` (own stream and work memory) and return at once; see frame_wait."""
        sd = (_f * 3)(*[float(x) for x in sun_dir])
        sr = (_f * 3)(*[float(x) for x in sun_radiance])
        if sample_end is None:
            sample_end = max(1, params.samples_per_pixel)
        r = (C.c_int * 4)(*rect) if rect is not None else None
        self._ck(lib.dsrt_render_frame_async(self._h, slot, C.byref(camera), sd, sr, int(sun_enabled), C.byref(params),
                                             seed, sample_begin, sample_end, r, int(keep_accum)))
        self.width, self.height = params.img_width, params.img_height

    def readback_rgb8_async(self, slot: int, total_spp: int, gamma: float, out: np.ndarray):
        """Enqueue tone map + copy of the slot's frame into `out` (keep it alive until frame_wait)."""
        self._ck(lib.dsrt_readback_rgb8_async(self._h, slot, total_spp, gamma, out.ctypes.data, out.nbytes))

    def frame_wait(self, slot: int) -> FrameStats:
        fs = FrameStats()
        self._ck(lib.dsrt_frame_wait(self._h, slot, C.byref(fs)))
        return fs

    def primary_hits(self, camera: Camera, params: RenderParams, sample=0, seed=DEFAULT_SEED):
        """First hits of one sample taken from the pipeline's own k_tile_entries + k_primary:
        (tri, t, u, v, leaf_node, leaf_record), each W*H with index x + y*W (y = 0 bottom row)."""
        n = params.img_width * params.img_height
        tri, node, rec = (np.empty(n, np.int32) for _ in range(3))
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        self._ck(lib.dsrt_primary_hits(self._h, C.byref(camera), C.byref(params), seed, sample, tri.ctypes.data,
                                       t.ctypes.data, u.ctypes.data, v.ctypes.data, node.ctypes.data, rec.ctypes.data))
        return tri, t, u, v, node, rec

    def frame_stats(self) -> FrameStats:
        fs = FrameStats()
        self._ck(lib.dsrt_frame_stats(self._h, C.byref(fs)))
        return fs

    def set_stage_timing(self, on: bool):
        self._ck(lib.dsrt_set_stage_timing(self._h, int(on)))

    def set_samples_per_batch(self, n: int):
        self._ck(lib.dsrt_set_samples_per_batch(self._h, int(n)))

    def set_bvh_builder(self, name: str):
        """'ploc' (default) or 'lbvh': the binary-tree stage used by the next upload_scene."""
        self._ck(lib.dsrt_set_bvh_builder(self._h, {"ploc": 0, "lbvh": 1}[name]))

    def readback_rgb8(self, total_spp: int, gamma: float = 2.0, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 3), dtype=np.uint8)
        self._ck(lib.dsrt_readback_rgb8(self._h, total_spp, gamma, out.ctypes.data, out.nbytes))
        return out

    def readback_accum(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 3), dtype=np.float32)
        self._ck(lib.dsrt_readback_accum(self._h, out.ctypes.data, out.nbytes))
        return out

    def accum_device(self):
        p, n = _vp(), C.c_size_t()
        self._ck(lib.dsrt_accum_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def upload_accum(self, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float32)
        self._ck(lib.dsrt_upload_accum(self._h, a.ctypes.data, a.nbytes))

    def trace_closest(self, rays: np.ndarray):
        """rays: (n, 8) float32 {ox,oy,oz,tmin,dx,dy,dz,tmax} -> (tri i32, t, u, v)."""
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = len(r)
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        self._ck(lib.dsrt_trace_closest(self._h, r.ctypes.data, n, tri.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data))
        return tri, t, u, v

    def trace_occluded(self, rays: np.ndarray) -> np.ndarray:
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        occ = np.empty(len(r), np.uint8)
        self._ck(lib.dsrt_trace_occluded(self._h, r.ctypes.data, len(r), occ.ctypes.data))
        return occ

    def trace_closest_device(self, rays_ptr: int, n: int, hits_ptr: int):
        self._ck(lib.dsrt_trace_closest_device(self._h, rays_ptr, n, hits_ptr))

    def trace_occluded_device(self, rays_ptr: int, n: int, occ_ptr: int):
        self._ck(lib.dsrt_trace_occluded_device(self._h, rays_ptr, n, occ_ptr))

    def first_hits(self, camera: Camera, width: int, height: int, jx=0.5, jy=0.5):
        n = width * height
        tri = np.empty(n, np.int32)
        t, u, v = (np.empty(n, np.float32) for _ in range(3))
        self._ck(lib.dsrt_first_hits(self._h, C.byref(camera), width, height, jx, jy,
                                     tri.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data))
        return tri, t, u, v
