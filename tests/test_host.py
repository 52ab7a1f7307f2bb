"""Host side of the path and the C-ABI surface (no GPU)."""
import ctypes
import hashlib
import json
import re
import zlib
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, POSE_FILE, ROOT, bits, cam12_of


def test_library_exports_every_declared_symbol(dsrt):
    declared = set()
    for h in ("dsrt.h", "dsrt_legacy.h"):
        text = (ROOT / "include" / h).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b((?:dsrt|gpu)_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(dsrt.lib, name), f"libdsrt.so does not export {name}"
    assert declared == set(dsrt.SYMBOLS)


def test_version(dsrt):
    assert "sm_100a" in dsrt.version()


def test_no_gpu_fails_loudly_or_creates(dsrt):
    """Without a CUDA device dsrt_create must fail with a message -- there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(dsrt.DsrtError, match="no usable CUDA device"):
        dsrt.Context(0)


def test_struct_sizes_match_reference_layout(dsrt):
    assert dsrt.triangle_dtype.itemsize == 116     # GPUTriangle, SURVEY 6
    assert dsrt.material_dtype.itemsize == 48      # GPUMaterial
    assert ctypes.sizeof(dsrt.Camera) == 104       # GPUCamera
    assert ctypes.sizeof(dsrt.RenderParams) == 48  # GPURenderParams


def test_pose_file_count_and_skips(dsrt, tmp_path, poses):
    assert poses.shape == (99, 9)                  # 100 text lines, line 1 is a comment
    p = tmp_path / "poses.txt"
    p.write_text("# comment\n\n1 2 3 4 5 6 90.5 1 2\nbad line here\n 7 8 9 10 11 12 0 0 0 trailing\n#1 2 3 4 5 6 7 8 9\n1 2 3\n")
    got = dsrt.read_pose_file(p)
    assert got.shape == (2, 9)
    assert got[0, 6] == np.float32(90.5) and got[1, 0] == 7
    with pytest.raises(dsrt.DsrtError):
        dsrt.read_pose_file(tmp_path / "missing.txt")


def test_pose_and_camera_match_reference_golden(dsrt, poses):
    g = json.loads((GOLDEN / "pose_camera_golden.json").read_text())
    assert g["num_poses"] == 99
    assert g["poses_sha256"] == hashlib.sha256(poses.tobytes()).hexdigest()
    for rec in g["frames"]:
        cam, sun, sep = dsrt.pose_to_model_frame(poses[rec["frame"]])
        assert [f"{x:08x}" for x in bits(cam)] == rec["cam_in_model"]
        assert [f"{x:08x}" for x in bits(sun)] == rec["sun_dir_model"]
        assert sep.hex() == rec["sep_m"]
        for key, c in rec["cameras"].items():
            W, H = map(int, key.split("x"))
            cm = dsrt.camera_look_at_origin(cam, 40.0, W, H)
            assert [f"{x:08x}" for x in bits(cam12_of(cm))] == c["cam12"], (rec["frame"], key)
            uvw = np.array(cm.u.tuple() + cm.v.tuple() + cm.w.tuple(), np.float32)
            assert [f"{x:08x}" for x in bits(uvw)] == c["uvw"]


def test_survey_appendix_b_vectors(dsrt, poses):
    """The three frames quoted in SURVEY.md Appendix B."""
    cam, sun, sep = dsrt.pose_to_model_frame(poses[0])
    assert [f"{x:08x}" for x in bits(cam)] == ["bf39e280", "00000000", "44df5f2c"]
    assert [f"{x:08x}" for x in bits(sun)] == ["3ea7d602", "bf41a4e1", "3f10e62b"]
    assert abs(sep - 1786.974222840) < 1e-8
    cm = dsrt.camera_look_at_origin(cam, 40.0, 800, 450)
    assert [f"{x:08x}" for x in bits(cam12_of(cm))[3:]] == [
        "c49088d8", "c42299f2", "bef08000", "451088d8", "80000000", "3f708e7b", "00000000", "44a299f2", "00000000"]
    cam98, _, sep98 = dsrt.pose_to_model_frame(poses[98])
    assert [f"{x:08x}" for x in bits(cam98)] == ["b9980000", "00000000", "420ef53c"]
    assert abs(sep98 - 35.739486696) < 1e-8


def test_synth_mesh_is_deterministic_and_in_envelope(dsrt):
    t1, m1 = dsrt.synth_mesh(10000)
    t2, m2 = dsrt.synth_mesh(10000)
    assert t1.tobytes() == t2.tobytes() and m1.tobytes() == m2.tobytes()
    v = np.stack([np.stack([t1[n][k] for k in "xyz"], 1) for n in ("v0", "v1", "v2")], 1)
    assert np.isfinite(v).all()
    lo, hi = v.min((0, 1)), v.max((0, 1))
    assert (lo >= [-950, -420, -25.5]).all() and (hi <= [950, 420, 25.5]).all(), (lo, hi)
    n = np.stack([t1["n0"][k] for k in "xyz"], 1)
    assert np.isfinite(n).all() and np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-5)
    assert np.array_equal(t1["n0"], t1["n1"]) and np.array_equal(t1["n0"], t1["n2"])   # flat normals
    # no axis-aligned (zero-thickness) triangle boxes: the reference cannot see those
    assert ((v.max(1) - v.min(1)) > 0).all()
    assert t1["material_id"].min() >= 0 and t1["material_id"].max() < len(m1)
    types = m1["type"][t1["material_id"]]
    frac = np.bincount(types, minlength=3) / len(types)
    assert frac[0] > 0.5 and frac[1] > 0.05 and frac[2] > 0.03
    td, md = dsrt.synth_mesh(5000, diffuse_only=True)
    assert (md["type"] == 0).all()
    for n_ in (0, 1, 2, 3, 17, 12345):
        tt, _ = dsrt.synth_mesh(n_)
        assert len(tt) == n_


def test_ppm_and_png_writers(dsrt, tmp_path):
    W, H = 13, 7
    rgb = (np.arange(W * H * 3) % 251).astype(np.uint8).reshape(H, W, 3)
    dsrt.write_ppm(tmp_path / "a.ppm", rgb, W, H)
    raw = (tmp_path / "a.ppm").read_bytes()
    assert raw == b"P6\n13 7\n255\n" + rgb.tobytes()            # gpu_render.cu:1105-1106
    dsrt.write_png(tmp_path / "a.png", rgb, W, H)
    png = (tmp_path / "a.png").read_bytes()
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    # decode by hand: IHDR + one IDAT
    pos, idat = 8, b""
    while pos < len(png):
        ln = int.from_bytes(png[pos:pos + 4], "big")
        typ = png[pos + 4:pos + 8]
        data = png[pos + 8:pos + 8 + ln]
        crc = int.from_bytes(png[pos + 8 + ln:pos + 12 + ln], "big")
        assert zlib.crc32(typ + data) == crc
        if typ == b"IHDR":
            assert int.from_bytes(data[:4], "big") == W and int.from_bytes(data[4:8], "big") == H and data[8:10] == b"\x08\x02"
        if typ == b"IDAT":
            idat += data
        pos += 12 + ln
    rows = zlib.decompress(idat)
    got = np.frombuffer(rows, np.uint8).reshape(H, 1 + W * 3)
    assert (got[:, 0] == 0).all() and np.array_equal(got[:, 1:].reshape(H, W, 3), rgb)
    try:
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(tmp_path / "a.png")), rgb)
    except ImportError:
        pass


def test_product_never_touches_the_oracle():
    """No file of the product path may reference oracle/ (the judge checks exactly this)."""
    for p in (ROOT / "deep-space-ray-tracer_b200").rglob("*"):
        if p.suffix in (".cu", ".cuh", ".h", ".cpp", ".py") or p.name == "Makefile":
            text = p.read_text(errors="ignore")
            assert "liboracle" not in text and "oracle/" not in text and "import oracle" not in text, p


def test_bench_touches_the_oracle_only_in_its_baseline_legs():
    """bench.py may execute oracle/ code (tests/oracle.py, refcpu.py, refgpu.py, synthmesh.py) only in the legs that
    time the reference: cpu_baseline, ref_gpu_kernel and run_reference (--impl reference).  The timed product path
    -- run_dsrt and everything it calls -- must not import them, and the reference arm must not import the product."""
    import ast
    tree = ast.parse((ROOT / "bench.py").read_text())
    checker_modules = {"oracle", "refcpu", "refgpu", "synthmesh"}
    allowed = {"cpu_baseline", "ref_gpu_kernel", "run_reference"}
    seen = set()
    for fn in [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))]:
        for node in ast.walk(fn):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name.split(".")[0] for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.module:
                names = [node.module.split(".")[0]]
            for nm in names:
                if nm in checker_modules:
                    assert fn.name in allowed, f"bench.py: {fn.name} imports {nm}"
                    seen.add(fn.name)
                if nm in ("dsrt_b200",):
                    assert fn.name not in ("run_reference", "cpu_baseline", "ref_gpu_kernel"), f"bench.py: {fn.name} maps the product library"
    for node in tree.body:                      # nothing at module level either
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            mods = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""]
            assert not (set(m.split(".")[0] for m in mods) & (checker_modules | {"dsrt_b200"})), mods
    assert seen == allowed


def test_box_screen_bounds_are_conservative(dsrt, poses):
    """dsrt_box_screen_bounds (host arithmetic behind k_primary's early miss): every point of the box,
    projected with the reference camera model u = (x + jx) / (W - 1) (gpu_render.cu:952-967), must
    land on a pixel inside the returned rectangle; and when a corner is behind the camera the
    function must say "no bound"."""
    import ctypes as C
    lib = dsrt.lib
    lib.dsrt_box_screen_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.dsrt_box_screen_bounds.restype = C.c_int
    rng = np.random.default_rng(3)
    W, H = 640, 360
    lo = np.array([-950.0, -420.0, -25.0], np.float32)
    hi = np.array([950.0, 420.0, 25.0], np.float32)
    views = [dsrt.pose_to_model_frame(poses[k])[0] for k in (0, 49, 98)] + [np.array(v, np.float32) for v in
             ((300.0, 120.0, 160.0), (2500.0, 900.0, 700.0), (-1400.0, 30.0, 12.0), (0.0, 0.5, 3.0))]
    bounded = 0
    for lookfrom in views:
        cam = dsrt.camera_look_at_origin(np.asarray(lookfrom, np.float32), 40.0, W, H)
        out = (C.c_int * 4)()
        ok = lib.dsrt_box_screen_bounds(C.byref(cam), lo.ctypes.data_as(C.POINTER(C.c_float)), hi.ctypes.data_as(C.POINTER(C.c_float)), W, H, out)
        x0, y0, x1, y1 = out[:]
        assert 0 <= x0 <= x1 <= W - 1 and 0 <= y0 <= y1 <= H - 1
        o = np.array(cam.origin.tuple(), np.float64)
        a = np.array(cam.lower_left_corner.tuple(), np.float64) - o
        hvec = np.array(cam.horizontal.tuple(), np.float64)
        vvec = np.array(cam.vertical.tuple(), np.float64)
        M = np.stack([a, hvec, vvec], 1)
        pts = rng.uniform(lo, hi, size=(20000, 3)).astype(np.float64)
        corners = np.array([[(hi if (c >> k) & 1 else lo)[k] for k in range(3)] for c in range(8)], np.float64)
        sol = np.linalg.solve(M, (np.concatenate([pts, corners]) - o).T).T          # (lambda, lambda u, lambda v)
        lam = sol[:, 0]
        if not ok:
            assert (x0, y0, x1, y1) == (0, 0, W - 1, H - 1)
            assert lam[-8:].min() <= 1e-6, "bounds were dropped although every corner is in front of the camera"
            continue
        bounded += 1
        assert lam.min() > 0
        px = np.floor(sol[:, 1] / lam * (W - 1))          # the pixel whose footprint [px, px + 1) holds the point
        py = np.floor(sol[:, 2] / lam * (H - 1))
        vis = (px >= 0) & (px <= W - 1) & (py >= 0) & (py <= H - 1)
        assert vis.any()
        assert (px[vis] >= x0).all() and (px[vis] <= x1).all() and (py[vis] >= y0).all() and (py[vis] <= y1).all()
    assert bounded >= 4


def test_cli_host_side_without_a_gpu(dsrt, tmp_path):
    """The ray_tracer CLI up to the point where it needs the device: the reference's flags and
    banner (src/main.cpp:194-215, 319-357), unknown flags ignored, the output directory emptied
    (:41-50), poses parsed and transformed in double precision -- and, with no CUDA device, a loud
    failure instead of any CPU rendering.  (On a GPU box the same command simply renders.)"""
    import subprocess
    import torch
    exe = ROOT / "deep-space-ray-tracer_b200" / "ray_tracer"
    if not exe.exists():
        pytest.skip("ray_tracer not built")
    out = tmp_path / "frames"
    out.mkdir()
    (out / "stale.txt").write_text("left over from an earlier run")
    r = subprocess.run([str(exe), "--input_txt", str(POSE_FILE), "--output_dir", str(out), "--synthetic_tris", "500",
                        "--frames", "0:3", "--width", "64", "--height", "36", "--spp", "2", "--max_depth", "3",
                        "--some_unknown_flag", "7"], capture_output=True, text=True, timeout=120)
    assert "Using output_dir: " + str(out) in r.stdout
    assert "Loaded 99 poses." in r.stdout
    assert r.stdout.count("=== Frame ") == 3
    assert "sep(cam, model) = 1786.97 m" in r.stdout                      # SURVEY appendix B, frame 0
    assert "sun_dir_model   = (0.327805, -0.756422, 0.566012)" in r.stdout
    assert not (out / "stale.txt").exists()
    if torch.cuda.is_available():
        assert r.returncode == 0 and (out / "frame_0000.ppm").exists() and (out / "frame_0002.png").exists()
    else:
        assert r.returncode != 0
        assert "no usable CUDA device" in r.stderr and "no CPU path" in r.stderr
        assert list(out.iterdir()) == []


def test_reference_arm_of_bench_does_not_map_the_product_library(tmp_path):
    """`bench.py --impl reference` times the reference's CPU classes; mesh, poses and camera come from oracle-side
    helpers, so libdsrt.so must not even be loaded by that process (the dynamic loader's own file log says so)."""
    import json
    import os
    import subprocess
    import sys
    env = dict(os.environ, LD_DEBUG="files", LD_DEBUG_OUTPUT=str(tmp_path / "ld"))
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--tris", "3000",
                        "--width", "96", "--height", "54", "--spp", "4", "--depth", "4", "--cpu_step", "2"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-1500:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port")
    loaded = "".join(f.read_text(errors="ignore") for f in tmp_path.glob("ld*"))
    assert "liboracle" in loaded or "libref_cpu" in loaded
    assert "libdsrt" not in loaded


def test_cli_resume_skips_frames_that_exist(tmp_path):
    """--resume keeps the output directory and renders only the frames whose files are missing; with every frame
    present nothing is left to do and no device is needed."""
    import subprocess
    exe = ROOT / "deep-space-ray-tracer_b200" / "ray_tracer"
    if not exe.exists():
        pytest.skip("ray_tracer not built")
    out = tmp_path / "frames"
    out.mkdir()
    for k in range(3):
        (out / f"frame_{k:04d}.ppm").write_bytes(b"P6\n64 36\n255\n" + bytes(64 * 36 * 3))
        (out / f"frame_{k:04d}.png").write_bytes(b"x")
    (out / "keep.txt").write_text("not wiped")
    r = subprocess.run([str(exe), "--input_txt", str(POSE_FILE), "--output_dir", str(out), "--synthetic_tris", "500", "--frames", "0:3",
                        "--width", "64", "--height", "36", "--spp", "2", "--max_depth", "3", "--resume"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("already rendered (--resume)") == 3 and "Nothing to render." in r.stdout
    assert (out / "keep.txt").exists() and (out / "frame_0001.png").read_bytes() == b"x"


def test_headers_are_plain_c_and_link_from_c(tmp_path):
    """include/*.h compiled by a C99 compiler with -pedantic (no C++, no CUDA types), a C program linked against
    libdsrt.so: record sizes, the host-side pose/camera entry points (Appendix B values) and, without a GPU, the loud
    failure of dsrt_create."""
    import subprocess
    import torch
    exe = tmp_path / "abi_check"
    lib_dir = ROOT / "deep-space-ray-tracer_b200"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "c" / "abi_check.c"),
           "-o", str(exe), f"-L{lib_dir}", "-ldsrt", f"-Wl,-rpath,{lib_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(POSE_FILE)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sm_100a" in r.stdout and "poses 99 sep 1786.974223" in r.stdout
    assert "cam -0.726112366 0 1786.97412 llc -1156.27637" in r.stdout          # SURVEY appendix B, frame 0
    if torch.cuda.is_available():
        assert "device ok" in r.stdout
    else:
        assert "no usable CUDA device" in r.stdout
