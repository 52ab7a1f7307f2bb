"""How much of a C2 / C5 frame is the sun shadow rays?  (GPU box; not a pytest file.)  Renders the frame with and
without the sun: the bounce rays are the same (the sun only adds radiance), so the difference is the any-hit rays."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import dsrt_b200 as d

poses = d.read_pose_file(ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt")
for tris_n, diffuse, spp, depth in ((1_000_000, False, 64, 10), (10_000_000, True, 256, 16)):
    tris, mats = d.synth_mesh(tris_n, diffuse_only=diffuse)
    with d.Context(0) as c:
        c.upload_scene(tris, mats)
        c.set_stage_timing(True)
        cam_pos, sun, _ = d.pose_to_model_frame(poses[0])
        cam = d.camera_look_at_origin(cam_pos, 40.0, 1920, 1080, spp, depth)
        p = d.make_params(1920, 1080, spp, depth)
        for sun_on in (True, False, True, False):
            fs = c.render(cam, sun, p, sun_enabled=sun_on)
            print(f"tris {tris_n} sun {sun_on}: frame {fs.device_ms:.2f} ms, traversal {fs.extend_ms:.2f} (primary {fs.primary_ms:.2f}), shade {fs.shade_ms:.2f}; "
                  f"rays primary {fs.primary_rays - fs.culled_primary_rays} bounce {fs.bounce_rays} shadow {fs.shadow_rays}", flush=True)
