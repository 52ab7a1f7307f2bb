"""Where the wall time of a short CLI run goes: context creation, upload + BVH build, the first
frame (work-memory allocation) and steady-state frames; and the batch-size sweep.  Not a pytest file."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import dsrt_b200 as d  # noqa: E402

out = {}
t0 = time.perf_counter()
tris, mats = d.synth_mesh(1_000_000)
out["synth_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
ctx = d.Context(0)
out["create_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
st = ctx.upload_scene(tris, mats)
out["upload_build_s"] = time.perf_counter() - t0
out["build_ms"] = st.build_ms
poses = d.read_pose_file(ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt")
cam_pos, sun, _ = d.pose_to_model_frame(poses[0])
W, H, SPP, DEPTH = 1920, 1080, 250, 50
cam = d.camera_look_at_origin(cam_pos, 40.0, W, H, SPP, DEPTH)
params = d.make_params(W, H, SPP, DEPTH)
rgb = np.empty((H, W, 3), np.uint8)
walls = []
for i in range(4):
    t0 = time.perf_counter()
    fs = ctx.render(cam, sun, params)
    ctx.readback_rgb8(SPP, 2.0, rgb)
    walls.append((time.perf_counter() - t0, fs.device_ms / 1e3, fs.batches))
out["frames_wall_device_batches"] = walls
sweep = {}
for spb in (4, 8, 16, 32, 64, 125, 250):
    ctx.set_samples_per_batch(spb)
    ctx.render(cam, sun, params)
    t0 = time.perf_counter()
    fs = ctx.render(cam, sun, params)
    sweep[spb] = {"wall_ms": 1e3 * (time.perf_counter() - t0), "device_ms": fs.device_ms, "mrays_s": fs.rays / fs.device_ms / 1e3,
                  "launches": fs.kernel_launches}
out["spb_sweep_250spp_depth50"] = sweep
t0 = time.perf_counter()
ctx.close()
out["destroy_s"] = time.perf_counter() - t0
print(json.dumps(out, indent=1))
