"""Close-range frame (pose 98) and far frame (pose 0) under the builder's environment switches (GPU box; not a pytest file)."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import dsrt_b200 as d

poses = d.read_pose_file(ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt")
tris, mats = d.synth_mesh(1_000_000)
for setting in sys.argv[1:] or [""]:
    for kv in setting.split(","):
        if "=" in kv:
            k, v = kv.split("=")
            os.environ[k] = v
    with d.Context(0) as c:
        st = c.upload_scene(tris, mats)
        c.set_stage_timing(True)
        out = []
        for k, spp, depth in ((0, 64, 10), (98, 64, 10)):
            cam_pos, sun, sep = d.pose_to_model_frame(poses[k])
            cam = d.camera_look_at_origin(cam_pos, 40.0, 1920, 1080, spp, depth)
            p = d.make_params(1920, 1080, spp, depth)
            c.render(cam, sun, p)
            fs = c.render(cam, sun, p)
            out.append(f"pose {k}: {fs.device_ms:.2f} ms (primary {fs.primary_ms:.2f}, trace {fs.extend_ms - fs.primary_ms:.2f})")
        print(f"{setting or 'default':40s} nodes {st.num_wide_nodes} sah {st.sah_cost:.2f} build {st.build_ms:.1f} ms | " + " | ".join(out), flush=True)
    for kv in setting.split(","):
        if "=" in kv:
            os.environ.pop(kv.split("=")[0], None)
