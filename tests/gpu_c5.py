"""BASELINE config 5 on the GPU box: 10M triangles, diffuse only, depth 16, 1080p, 256 spp
(traversal / memory-bound case).  Not a pytest file; prints one JSON line."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import dsrt_b200 as d  # noqa: E402

tris_n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 256
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 2
poses = d.read_pose_file(ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt")
cam_pos, sun, _ = d.pose_to_model_frame(poses[0])
t0 = time.time()
tris, mats = d.synth_mesh(tris_n, diffuse_only=True)
gen_s = time.time() - t0
ctx = d.Context(0)
st = ctx.upload_scene(tris, mats)
W, H, DEPTH = 1920, 1080, 16
cam = d.camera_look_at_origin(cam_pos, 40.0, W, H, spp, DEPTH)
params = d.make_params(W, H, spp, DEPTH)
ctx.set_stage_timing(True)
out = {"tris": tris_n, "gen_s": gen_s, "nodes": st.num_wide_nodes, "levels": st.num_levels, "build_ms": st.build_ms,
       "upload_ms": st.upload_ms, "sah": st.sah_cost, "runs": []}
for i in range(frames):
    fs = ctx.render(cam, sun, params)
    out["runs"].append(dict(device_ms=fs.device_ms, rays=int(fs.rays), primary=int(fs.primary_rays), bounce=int(fs.bounce_rays),
                            shadow=int(fs.shadow_rays), mrays_s=fs.rays / fs.device_ms / 1e3, traversal_ms=fs.extend_ms,
                            shade_ms=fs.shade_ms, launches=fs.kernel_launches, batches=fs.batches))
bpr = 48 + 80 * 8 + 192          # B(10M) = 880 B/ray
r = out["runs"][-1]
out["roofline_frac_of_6549"] = (r["rays"] * bpr / (r["traversal_ms"] * 1e-3) / 1e9) / 6549.1
print(json.dumps(out))
