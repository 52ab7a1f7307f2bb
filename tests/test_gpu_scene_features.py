"""GPUScene semantics beyond the bare triangle mesh, through the C ABI on the GPU: spheres with
the emissive-sphere light/BRDF mixture and textured OBJ scenes (f1, f3).

Checkers: the CPU oracle (oracle/oracle.c), the unmodified reference kernel when
oracle/_ref/libref_gpu.so travelled to the box, and the golden fixture
tests/golden/ref_gpu_features.npz produced from that kernel by `oracle/gen_golden.py --gpu`.
Bars: hits bit-exact, LCG-replay images byte-identical to the reference kernel's."""
from pathlib import Path

import numpy as np
import pytest

import oracle
import refgpu
from conftest import bits, cam12_of
from gpu_probe import random_rays

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
OBJ_DIR = ROOT / "tests" / "golden" / "obj"
W, H, SPP, DEPTH = 160, 90, 8, 8
SUN = np.array([0.327804625, -0.756422102, 0.566012084], np.float32)      # sun_dir_model of the shipped poses


def sphere_scene(dsrt, with_lights=True, ntris=3000):
    """The synthetic station plus a handful of spheres of every material class in front of it."""
    tris, mats = dsrt.synth_mesh(ntris)
    extra = np.zeros(5, dsrt.material_dtype)
    extra["albedo_tex"] = -1
    extra["type"] = [dsrt.MAT_LAMBERTIAN, dsrt.MAT_METAL, dsrt.MAT_DIELECTRIC, dsrt.MAT_DIFFUSE_LIGHT, dsrt.MAT_DIFFUSE_LIGHT]
    for k, a in enumerate([(0.7, 0.3, 0.2), (0.9, 0.85, 0.6), (1, 1, 1), (1, 1, 1), (1, 1, 1)]):
        extra["albedo"][k] = a
    extra["fuzz"] = [0, 0.15, 0, 0, 0]
    extra["ref_idx"] = [1.5, 1.5, 1.45, 1, 1]
    extra["emissive"][3] = (0.9, 0.8, 0.7) if with_lights else (0, 0, 0)
    extra["emissive"][4] = (0.0, 0.4, 0.9) if with_lights else (0, 0, 0)
    base = len(mats)
    mats = np.concatenate([mats, extra])
    sph = np.zeros(6, dsrt.sphere_dtype)
    centers = [(-420, 60, 180), (-120, -150, 140), (200, 120, 220), (480, -60, 260), (60, 330, 320), (0, 0, 90)]
    radii = [110, 90, 100, 70, 55, 40]
    mids = [base + 0, base + 1, base + 2, base + 3, base + 4, base + 0]
    for k in range(6):
        sph["center"][k] = centers[k]
        sph["radius"][k] = radii[k]
        sph["material_id"][k] = mids[k]
    return tris, mats, sph


def camera_at(dsrt, lookfrom, w=W, h=H, spp=SPP, depth=DEPTH):
    return dsrt.camera_look_at_origin(np.array(lookfrom, np.float32), 40.0, w, h, spp, depth)


def hits_equal(ours, ref, min_hits, ref_marks_spheres=True):
    tri, t, u, v = ours
    rtri, rt, ru, rv = ref
    if not ref_marks_spheres:                      # the reference's HitRecord has tri_index = -1 for spheres
        tri = np.where(tri < -1, -1, tri)
        hit = bits(rt) != bits(np.float32(1e9))
    else:
        hit = rtri != -1
    same = tri == rtri
    tie = (~same) & (bits(t) == bits(rt))          # equal-t ties may pick the other primitive (SURVEY A-8)
    assert (same | tie).all(), f"{(~(same | tie)).sum()} hits differ"
    assert hit.sum() >= min_hits
    assert np.array_equal(bits(t)[hit & same], bits(rt)[hit & same])
    assert np.array_equal(bits(u)[hit & same], bits(ru)[hit & same]) and np.array_equal(bits(v)[hit & same], bits(rv)[hit & same])


def test_spheres_hits_bit_exact_vs_oracle(dsrt):
    tris, mats, sph = sphere_scene(dsrt)
    cam = camera_at(dsrt, (350, 260, 1500))
    sc = oracle.Scene(tris, mats, spheres=sph)
    with dsrt.Context(0) as c:
        c.upload_scene(tris, mats, spheres=sph)
        ours = c.first_hits(cam, W, H, 0.5, 0.5)
        ref = sc.first_hits(cam12_of(cam), W, H, 0.5, 0.5)
        hits_equal(ours, ref, 2000)
        assert (ours[0] < -1).sum() > 300 and len(np.unique(ours[0][ours[0] < -1])) >= 5     # spheres are seen
        st = c.scene_stats()
        rays = random_rays(60000, 9, np.array(st.scene_lo[:]) - 100, np.array(st.scene_hi[:]) + 300)
        hits_equal(c.trace_closest(rays), sc.trace(rays), 15000)
        occ = c.trace_occluded(rays)
        assert np.array_equal(occ, sc.occluded(rays))
        assert np.array_equal(occ.astype(bool), c.trace_closest(rays)[0] != -1)


def test_sphere_only_scene(dsrt):
    tris, mats, sph = sphere_scene(dsrt, ntris=0)
    cam = camera_at(dsrt, (100, 80, 1200), 96, 54, 4, 6)
    sc = oracle.Scene(tris, mats, spheres=sph)
    with dsrt.Context(0) as c:
        st = c.upload_scene(tris, mats, spheres=sph)
        assert st.num_triangles == 0 and st.num_wide_nodes == 0
        hits_equal(c.first_hits(cam, 96, 54, 0.5, 0.5), sc.first_hits(cam12_of(cam), 96, 54, 0.5, 0.5), 300)
        fs = c.render(cam, SUN, dsrt.make_params(96, 54, 4, 6, dsrt.RNG_REF_LCG))
        img = c.readback_rgb8(4, 2.0)
        _, orgb, counts = sc.render(cam12_of(cam), 96, 54, 4, 6, SUN, rng_mode=1)
        assert fs.primary_rays == counts[0] and (img == orgb).all(2).mean() >= 0.99 and img.max() > 0


@pytest.mark.parametrize("with_lights", [True, False])
def test_spheres_images_vs_oracle(dsrt, with_lights):
    """LCG replay and Philox against the oracle's restatement (CPU libm differs in the last place
    of sinf/cosf/powf: >= 99 % byte-identical pixels, mean |difference| of the float sums small)."""
    tris, mats, sph = sphere_scene(dsrt, with_lights)
    cam = camera_at(dsrt, (350, 260, 1500))
    sc = oracle.Scene(tris, mats, spheres=sph)
    with dsrt.Context(0) as c:
        c.upload_scene(tris, mats, spheres=sph)
        for mode in (dsrt.RNG_REF_LCG, dsrt.RNG_PHILOX):
            fs = c.render(cam, SUN, dsrt.make_params(W, H, SPP, DEPTH, mode))
            img, acc = c.readback_rgb8(SPP, 2.0), c.readback_accum()
            oacc, orgb, counts = sc.render(cam12_of(cam), W, H, SPP, DEPTH, SUN, rng_mode=1 if mode == dsrt.RNG_REF_LCG else 0)
            assert (img == orgb).all(2).mean() >= 0.99, mode
            assert np.abs(acc - oacc).mean() < 0.01 * SPP
            assert fs.primary_rays == counts[0]
            assert abs(int(fs.bounce_rays) - int(counts[1])) <= 0.01 * counts[1] + 5
            assert abs(int(fs.shadow_rays) - int(counts[2])) <= 0.01 * counts[2] + 5


@pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref/libref_gpu.so not present")
@pytest.mark.parametrize("with_lights", [True, False])
def test_spheres_identical_to_live_reference_kernel(dsrt, with_lights):
    tris, mats, sph = sphere_scene(dsrt, with_lights)
    lookfrom = np.array((350, 260, 1500), np.float32)
    cam = camera_at(dsrt, lookfrom)
    ref = refgpu.RefScene(tris, mats, lookfrom, SUN, W, H, SPP, DEPTH, spheres=sph)
    assert ref.normal_mismatches == 0 and np.array_equal(bits(cam12_of(cam)), bits(ref.camera12))
    a = ref.arrays()
    assert a["counts"][2] == len(sph) and np.array_equal(a["spheres"], sph.view(np.uint32).reshape(-1, 6))
    with dsrt.Context(0) as c:
        c.upload_scene(tris, mats, spheres=sph)
        hits_equal(c.first_hits(cam, W, H, 0.3, 0.6), ref.first_hits(0.3, 0.6), 2000, ref_marks_spheres=False)
        c.render(cam, SUN, dsrt.make_params(W, H, SPP, DEPTH, dsrt.RNG_REF_LCG))
        img = c.readback_rgb8(SPP, 2.0)
    rimg, _ = ref.render()
    ref.close()
    assert np.array_equal(img, rimg), f"{(img != rimg).any(2).sum()} pixels differ from the reference kernel"
    assert (img.reshape(-1, 3).max(1) > 0).mean() > 0.02


@pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref/libref_gpu.so not present")
def test_reference_gpuscene_with_spheres_through_the_legacy_upload(dsrt):
    """What the `gpu_render_scene` shim does with a reference GPUScene that holds spheres (round 1 refused them):
    triangles, materials and spheres are taken from the DEVICE arrays the reference's build_gpu_scene produced
    (dsrt_scene_upload_legacy); the LCG frame is the reference kernel's, byte for byte.  And the device path
    range-checks material ids with a kernel."""
    import torch
    tris, mats, sph = sphere_scene(dsrt, True)
    lookfrom = np.array((350, 260, 1500), np.float32)
    cam = camera_at(dsrt, lookfrom)
    ref = refgpu.RefScene(tris, mats, lookfrom, SUN, W, H, SPP, DEPTH, spheres=sph)
    rimg, _ = ref.render()
    with dsrt.Context(0) as c:
        st, masked = c.upload_scene_legacy(ref.legacy_scene_ptr(), False)
        assert st.num_triangles == len(tris) and masked == 0
        c.render(cam, SUN, dsrt.make_params(W, H, SPP, DEPTH, dsrt.RNG_REF_LCG))
        img = c.readback_rgb8(SPP, 2.0)
        assert np.array_equal(img, rimg), f"{(img != rimg).any(2).sum()} pixels differ from the reference kernel"
        bad = tris.copy()
        bad["material_id"][17] = len(mats) + 3
        d_tris = torch.from_numpy(bad.view(np.uint8).copy()).cuda()
        d_mats = torch.from_numpy(mats.view(np.uint8).copy()).cuda()
        with pytest.raises(dsrt.DsrtError, match="material_id"):
            c.upload_scene_device(d_tris.data_ptr(), len(bad), d_mats.data_ptr(), len(mats))
        d_ok = torch.from_numpy(tris.view(np.uint8).copy()).cuda()
        d_sph = torch.from_numpy(sph.view(np.uint8).copy()).cuda()
        c.upload_scene_device(d_ok.data_ptr(), len(tris), d_mats.data_ptr(), len(mats), d_sph.data_ptr(), len(sph))
        c.render(cam, SUN, dsrt.make_params(W, H, SPP, DEPTH, dsrt.RNG_REF_LCG))
        assert np.array_equal(c.readback_rgb8(SPP, 2.0), rimg)
    ref.close()


def panel_obj(dsrt, tmp_path, jpeg=False):
    """A textured scene that exercises tex2D: a wall of quads whose uvs run outside [0,1] (wrap),
    two textures of different sizes, an untextured floor, a metal and a glass block.  jpeg: the two
    textures are JPEG files (baseline 4:2:0 and progressive 4:4:4), which the reference decodes with stb_image."""
    rng = np.random.default_rng(5)
    ta, tb = rng.integers(0, 256, (16, 24, 3)).astype(np.uint8), rng.integers(0, 256, (9, 7, 3)).astype(np.uint8)
    na, nb = ("a.jpg", "b.jpg") if jpeg else ("a.ppm", "b.png")
    if jpeg:
        from PIL import Image
        Image.fromarray(np.repeat(np.repeat(ta, 3, 0), 3, 1)).save(tmp_path / na, quality=88, subsampling=2)
        Image.fromarray(np.repeat(np.repeat(tb, 4, 0), 4, 1)).save(tmp_path / nb, quality=92, subsampling=0, progressive=True)
    else:
        dsrt.write_ppm(tmp_path / na, ta, 24, 16)
        dsrt.write_png(tmp_path / nb, tb, 7, 9)
    (tmp_path / "p.mtl").write_text(f"newmtl ta\nmap_Kd {na}\nnewmtl tb\nKd 0.2 0.2 0.2\nmap_Kd {nb}\nnewmtl floor\nKd 0.6 0.6 0.55\n"
                                    "newmtl steel\nKs 0.8 0.8 0.85\nNs 400\nnewmtl glass\nd 0.3\nNi 1.5\n")
    L = ["mtllib p.mtl"]
    nv = 0
    n = 12
    for j in range(n + 1):
        for i in range(n + 1):
            L.append(f"v {(i - n / 2) * 10 + rng.normal(0, 0.8):.5f} {(j - n / 2) * 10 + rng.normal(0, 0.8):.5f} {rng.normal(0, 1.5):.5f}")
            L.append(f"vt {i / n * 3 - 1:.5f} {j / n * 2.5 - 0.75:.5f}")
    nv = (n + 1) ** 2
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i + 1
            L.append("usemtl " + ("ta" if (i + j) % 2 else "tb"))
            L.append(f"f {a}/{a} {a + 1}/{a + 1} {a + n + 2}/{a + n + 2} {a + n + 1}/{a + n + 1}")
    L += ["v -80 -62.3 -40.1", "v 80.2 -61.7 -39.8", "v 79.6 -60.9 60.3", "v -80.4 -61.1 59.7", "usemtl floor",
          f"f {nv + 1} {nv + 2} {nv + 3} {nv + 4}"]
    L += ["v -30 -20 12.1", "v -10.3 -20.4 11.7", "v -20.1 0.2 13.2", "usemtl steel", f"f {nv + 5} {nv + 6} {nv + 7}",
          "v 10 -10 20.3", "v 31.2 -11 19.6", "v 21 11.4 21.1", "usemtl glass", f"f {nv + 8} {nv + 9} {nv + 10}"]
    (tmp_path / "p.obj").write_text("\n".join(L) + "\n")
    return tmp_path / "p.obj"


def obj_upload(dsrt, c, m):
    return c.upload_scene(m["tris"], m["mats"], textures=m["textures"], texture_pool=m["texture_pool"])


def test_textured_obj_vs_oracle(dsrt, tmp_path):
    m = dsrt.load_obj(panel_obj(dsrt, tmp_path))
    assert len(m["tris"]) == 2 * 144 + 2 + 2 and len(m["textures"]) == 2 and m["textures_flipped"]
    cam = camera_at(dsrt, (25, 30, 170))
    sun = np.array([-0.3, -0.5, -0.8], np.float32)       # lights the camera side of the wall (light comes from -sun_dir)
    sc = oracle.Scene(m["tris"], m["mats"], textures=m["textures"], texture_pool=m["texture_pool"])
    with dsrt.Context(0) as c:
        obj_upload(dsrt, c, m)
        hits_equal(c.first_hits(cam, W, H, 0.5, 0.5), sc.first_hits(cam12_of(cam), W, H, 0.5, 0.5), 3000)
        for mode in (dsrt.RNG_REF_LCG, dsrt.RNG_PHILOX):
            c.render(cam, sun, dsrt.make_params(W, H, SPP, DEPTH, mode), sun_radiance=(0.9, 0.8, 0.7))
            img, acc = c.readback_rgb8(SPP, 2.0), c.readback_accum()
            oacc, orgb, _ = sc.render(cam12_of(cam), W, H, SPP, DEPTH, sun, sun_rad=(0.9, 0.8, 0.7), rng_mode=1 if mode == dsrt.RNG_REF_LCG else 0)
            assert (img == orgb).all(2).mean() >= 0.99
            assert np.abs(acc - oacc).mean() < 0.01 * SPP
            # the texture shows: many distinct colours on the wall (a flat albedo would give few)
            assert len(np.unique(img.reshape(-1, 3), axis=0)) > 500


@pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref/libref_gpu.so not present")
@pytest.mark.parametrize("scene", ["panel", "panel_jpeg", "station"])
def test_obj_scene_identical_to_live_reference_kernel(dsrt, tmp_path, scene):
    """The reference loads the OBJ with its own triangle_mesh + stb_image and renders with its own
    kernel; this path loads it with dsrt_mesh_load_obj and renders through the C ABI."""
    if scene in ("panel", "panel_jpeg"):
        path, scale, lookfrom = panel_obj(dsrt, tmp_path, jpeg=scene == "panel_jpeg"), 1.0, np.array((25, 30, 170), np.float32)
    else:
        path, scale, lookfrom = OBJ_DIR / "station.obj", 2.5, np.array((3, 4, 22), np.float32)
    sun = np.array([-0.3, -0.5, -0.8], np.float32)
    m = dsrt.load_obj(path, scale)
    cam = camera_at(dsrt, lookfrom)
    ref = refgpu.RefScene(None, None, lookfrom, sun, W, H, SPP, DEPTH, obj=path, obj_scale=scale)
    a = ref.arrays()
    assert np.array_equal(a["tris"], m["tris"].view(np.uint32).reshape(-1, 29))
    assert np.array_equal(a["mats"], m["mats"].view(np.uint32).reshape(-1, 12))
    assert np.array_equal(a["textures"], m["textures"].view(np.int32).reshape(-1, 3))
    assert np.array_equal(a["pool"], m["texture_pool"].view(np.uint32))
    with dsrt.Context(0) as c:
        obj_upload(dsrt, c, m)
        hits_equal(c.first_hits(cam, W, H, 0.5, 0.5), ref.first_hits(0.5, 0.5), 100)
        c.render(cam, sun, dsrt.make_params(W, H, SPP, DEPTH, dsrt.RNG_REF_LCG))     # the reference's sun radiance (1e5 ...)
        img = c.readback_rgb8(SPP, 2.0)
    rimg, _ = ref.render()
    ref.close()
    assert np.array_equal(img, rimg), f"{(img != rimg).any(2).sum()} pixels differ from the reference kernel"


def test_features_vs_reference_golden(dsrt):
    """The same comparisons against the committed outputs of the reference kernel
    (tests/golden/ref_gpu_features.npz, oracle/gen_golden.py --gpu) -- holds without oracle/_ref."""
    p = ROOT / "tests" / "golden" / "ref_gpu_features.npz"
    if not p.exists():
        pytest.skip("tests/golden/ref_gpu_features.npz not generated yet")
    g = np.load(p)
    w, h, spp, depth = (int(g[k]) for k in ("W", "H", "SPP", "DEPTH"))
    for lights in (1, 0):
        tris, mats, sph = sphere_scene(dsrt, bool(lights))
        cam = camera_at(dsrt, (350, 260, 1500), w, h, spp, depth)
        with dsrt.Context(0) as c:
            c.upload_scene(tris, mats, spheres=sph)
            ours = c.first_hits(cam, w, h, 0.5, 0.5)
            hits_equal(ours, tuple(g[f"sph{lights}_first_{k}"] for k in ("tri", "t", "u", "v")), 2000, ref_marks_spheres=False)
            c.render(cam, SUN, dsrt.make_params(w, h, spp, depth, dsrt.RNG_REF_LCG))
            assert np.array_equal(c.readback_rgb8(spp, 2.0), g[f"sph{lights}_lcg_rgb"])
    m = dsrt.load_obj(OBJ_DIR / "station.obj", 2.5)
    cam = camera_at(dsrt, (3, 4, 22), w, h, spp, depth)
    with dsrt.Context(0) as c:
        obj_upload(dsrt, c, m)
        c.render(cam, np.array([-0.3, -0.5, -0.8], np.float32), dsrt.make_params(w, h, spp, depth, dsrt.RNG_REF_LCG))
        assert np.array_equal(c.readback_rgb8(spp, 2.0), g["station_lcg_rgb"])
