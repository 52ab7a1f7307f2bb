"""GPU probe (not a pytest file): the C2 frame with and without the sun's shadow rays.
    gpurun --timeout 600 -- 'python tests/gpu_sunsplit.py'
Measured on B200 (round 1): 132.7 M primary, 20.6 M bounce and only 0.79 M shadow rays at pose 0 (most
surfaces the camera sees face away from the light); the shadow rays cost 0.3 ms of the 13.2 ms frame."""
import sys; from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, dsrt_b200 as d
poses = d.read_pose_file(ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt")
cam_pos, sun, _ = d.pose_to_model_frame(poses[0])
tris, mats = d.synth_mesh(1000000)
W, H, SPP, DEPTH = 1920, 1080, 64, 10
cam = d.camera_look_at_origin(cam_pos, 40.0, W, H, SPP, DEPTH)
with d.Context(0) as ctx:
    ctx.upload_scene(tris, mats)
    ctx.set_stage_timing(True)
    p = d.make_params(W, H, SPP, DEPTH, d.RNG_PHILOX)
    for sun_on in (True, False, True, False):
        for _ in range(2):
            fs = ctx.render(cam, sun, p, sun_enabled=sun_on)
        print("sun", sun_on, "ms", round(fs.device_ms, 2), "trav", round(fs.extend_ms, 2), "primary", round(fs.primary_ms, 2), "shade", round(fs.shade_ms, 2),
              "rays p/b/s", fs.primary_rays, fs.bounce_rays, fs.shadow_rays)
