#!/bin/bash
# compute-sanitizer on the smallest case that runs every frame kernel (not a pytest file); ONE tool per gpurun call.
#   gpurun --timeout 900 -- 'bash tests/gpu_sanitize.sh memcheck|racecheck <tag>'
TOOL=${1:-memcheck}; TAG=${2:-san}
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import dsrt_b200 as d
poses = d.read_pose_file("tests/golden/rendezvous_1s_dt0_01s.txt")
tris, mats = d.synth_mesh(3000)
with d.Context(0) as c:
    c.upload_scene(tris, mats)
    for frame, (w, h, spp, depth, mode) in ((98, (64, 36, 4, 20, d.RNG_PHILOX)), (60, (48, 27, 2, 6, d.RNG_REF_LCG))):
        cam_pos, sun, _ = d.pose_to_model_frame(poses[frame])
        cam = d.camera_look_at_origin(cam_pos, 40.0, w, h, spp, depth)
        p = d.make_params(w, h, spp, depth, mode)
        fs = c.render(cam, sun, p)
        img = c.readback_rgb8(spp, 2.0)
        c.primary_hits(cam, p, sample=0)
        c.render_async(1, cam, sun, p); c.frame_wait(1)
        print(frame, fs.rays, int(img.sum()))
    rays = np.zeros((500, 8), np.float32); rays[:, 2] = 300; rays[:, 3] = 1e-3; rays[:, 6] = -1; rays[:, 7] = 1e9
    rays[:, 0] = np.linspace(-900, 900, 500)
    print(int((c.trace_closest(rays)[0] >= 0).sum()), int(c.trace_occluded(rays).sum()))
print("case done")
PY
python /tmp/san_case.py > gpurun_out/san_plain_$TAG.log 2>&1 && timeout 800 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_${TOOL}_$TAG.log 2>&1; echo "sanitizer_rc=$?"
tail -5 gpurun_out/san_plain_$TAG.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard" gpurun_out/sanitizer_${TOOL}_$TAG.log | head -20; tail -3 gpurun_out/sanitizer_${TOOL}_$TAG.log
