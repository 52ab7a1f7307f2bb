"""Where the time of a C3 frame goes as the camera closes in (GPU box; not a pytest file)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import dsrt_b200 as d

poses = d.read_pose_file(ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt")
tris, mats = d.synth_mesh(1_000_000)
with d.Context(0) as c:
    c.upload_scene(tris, mats)
    c.set_stage_timing(True)
    p = d.make_params(1920, 1080, 250, 50)
    for k in (0, 40, 70, 85, 92, 96, 98, 98):
        cam_pos, sun, sep = d.pose_to_model_frame(poses[k])
        cam = d.camera_look_at_origin(cam_pos, 40.0, 1920, 1080, 250, 50)
        fs = c.render(cam, sun, p)
        trav = fs.primary_rays - fs.culled_primary_rays
        print(f"pose {k:2d} sep {sep:7.1f} m: frame {fs.device_ms:7.2f} ms = primary {fs.primary_ms:6.2f} + trace/tail {fs.extend_ms - fs.primary_ms:6.2f} + shade {fs.shade_ms:6.2f} "
              f"+ other {fs.device_ms - fs.extend_ms - fs.shade_ms:5.2f}; rays primary {trav / 1e6:6.1f}M bounce {fs.bounce_rays / 1e6:6.1f}M shadow {fs.shadow_rays / 1e6:5.1f}M "
              f"-> {(trav + fs.bounce_rays + fs.shadow_rays) / fs.device_ms / 1e3:6.0f} Mrays/s traversed; launches {fs.kernel_launches}, batches {fs.batches}", flush=True)
