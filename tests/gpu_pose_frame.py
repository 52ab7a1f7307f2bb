"""One frame of the pose sequence at a given pose (GPU box; not a pytest file): the command the close-range ncu lists profile.

    python tests/gpu_pose_frame.py <pose index> [spp] [depth]
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import dsrt_b200 as d

k = int(sys.argv[1]) if len(sys.argv) > 1 else 98
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 10
poses = d.read_pose_file(ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt")
tris, mats = d.synth_mesh(1_000_000)
with d.Context(0) as c:
    c.upload_scene(tris, mats)
    c.set_stage_timing(True)
    cam_pos, sun, sep = d.pose_to_model_frame(poses[k])
    cam = d.camera_look_at_origin(cam_pos, 40.0, 1920, 1080, spp, depth)
    p = d.make_params(1920, 1080, spp, depth)
    for _ in range(2):
        fs = c.render(cam, sun, p)
    trav = fs.primary_rays - fs.culled_primary_rays
    print(f"pose {k} sep {sep:.1f} m, {spp} spp, depth {depth}: frame {fs.device_ms:.2f} ms = primary {fs.primary_ms:.2f} + trace/tail "
          f"{fs.extend_ms - fs.primary_ms:.2f} + shade {fs.shade_ms:.2f}; camera rays {trav / 1e6:.1f}M bounce {fs.bounce_rays / 1e6:.1f}M "
          f"shadow {fs.shadow_rays / 1e6:.1f}M", flush=True)
