#!/bin/bash
# Full ncu capture of the first two k_trace launches of a pose-98 frame (GPU box; not a pytest file); summarised on the box.
#   gpurun --timeout 900 -- 'bash tests/gpu_r2_close_ncu.sh <tag>'
TAG=${1:-cn}
mkdir -p gpurun_out
CMD="python tests/gpu_pose_frame.py 98 16 10"
$CMD | tee gpurun_out/pose98_frame_$TAG.txt
# skip the warm-up frame (k_primary, k_primary_lists, 10 k_trace) and the second frame's k_primary: capture its k_primary_lists and first k_trace
ncu --set full --clock-control none --import-source on -k 'regex:k_trace|k_primary' --launch-skip 13 -c 2 -f -o /tmp/prof_pose98_$TAG $CMD > gpurun_out/ncu_pose98full_$TAG.log 2>&1
echo ncu_rc=$?; tail -2 gpurun_out/ncu_pose98full_$TAG.log
python profiles/summarize_ncu.py kernel /tmp/prof_pose98_$TAG.ncu-rep > gpurun_out/pose98_kernel_$TAG.txt; echo k_rc=$?
python profiles/summarize_ncu.py lines /tmp/prof_pose98_$TAG.ncu-rep k_trace 0 45 > gpurun_out/pose98_lines_$TAG.txt; echo l_rc=$?
python profiles/summarize_ncu.py lines /tmp/prof_pose98_$TAG.ncu-rep k_primary_lists 0 30 >> gpurun_out/pose98_lines_$TAG.txt; echo l2_rc=$?
du -sh gpurun_out
