"""Generates tests/golden/jpeg: small JPEG files of every kind the texture decoder covers, and the RGB bytes
the REFERENCE decodes them to -- its vendored stb_image, `stbi_load(path, &w, &h, &n, 3)`
(src/gpu_scene_builder.cpp:214), called in oracle/_ref/libref_host.so (the reference's loader compiled from
/root/reference by oracle/Makefile).  Run in the container that has /root/reference:

    python tests/golden/gen_jpeg_golden.py

The .jpg files are written with PIL (only as a source of valid JPEG streams; seeds fixed); expected.npz holds
the reference's decode of each.  tests/test_obj_loader.py checks deep-space-ray-tracer_b200/host/jpeg_decode.cpp
against them byte for byte.
"""
import ctypes as C
from pathlib import Path

import numpy as np
from PIL import Image

HERE = Path(__file__).resolve().parent
OUT = HERE / "jpeg"
REF = HERE.parent.parent / "oracle" / "_ref" / "libref_host.so"


def ref_decode(lib, path):
    w, h, n = C.c_int(), C.c_int(), C.c_int()
    p = lib.stbi_load(str(path).encode(), w, h, n, 3)
    assert p, path
    return np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()


def main():
    lib = C.CDLL(str(REF))
    lib.stbi_load.restype = C.POINTER(C.c_ubyte)
    lib.stbi_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    lib.stbi_set_flip_vertically_on_load.argtypes = [C.c_int]
    lib.stbi_set_flip_vertically_on_load(0)
    OUT.mkdir(exist_ok=True)
    rng = np.random.default_rng(7)
    h, w = 21, 29
    smooth = np.clip(np.add.outer(np.linspace(10, 220, h), np.linspace(0, 60, w))[..., None] * np.array([1.0, 0.7, 0.4])
                     + rng.normal(0, 14, (h, w, 3)), 0, 255).astype(np.uint8)
    noise = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    files = {
        "base_444.jpg": (smooth, dict(quality=85, subsampling=0)),
        "base_422.jpg": (smooth, dict(quality=70, subsampling=1)),
        "base_420_noise.jpg": (noise, dict(quality=40, subsampling=2)),
        "base_420_restart.jpg": (smooth, dict(quality=75, subsampling=2, restart_marker_blocks=1)),
        "base_optimized.jpg": (smooth, dict(quality=60, subsampling=1, optimize=True)),
        "prog_444.jpg": (smooth, dict(quality=80, subsampling=0, progressive=True)),
        "prog_420_noise.jpg": (noise, dict(quality=55, subsampling=2, progressive=True)),
        "grey.jpg": (smooth[..., 0], dict(quality=80)),
        "grey_prog.jpg": (noise[..., 1], dict(quality=65, progressive=True)),
        "one_pixel.jpg": (smooth[:1, :1], dict(quality=75)),
    }
    expected = {}
    for name, (img, opts) in files.items():
        Image.fromarray(img).save(OUT / name, **opts)
        expected[name] = ref_decode(lib, OUT / name)
    np.savez_compressed(OUT / "expected.npz", **expected)
    print({k: v.shape for k, v in expected.items()})


if __name__ == "__main__":
    main()
