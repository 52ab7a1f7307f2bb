#!/bin/bash
# Close-range frames (GPU box; not a pytest file): the pose split under library variants, then the ncu launch list of one
# pose-98 frame (which kernel the camera rays' time goes to: k_primary or k_primary_lists).
#   gpurun --timeout 900 -- 'bash tests/gpu_r2_close.sh <tag> "<variant tags>"'
TAG=${1:-cl}; VARS=${2:-}
mkdir -p gpurun_out
for v in main $VARS; do
  LIB=""; [ "$v" != main ] && LIB=$PWD/deep-space-ray-tracer_b200/build/variants/libdsrt_$v.so
  echo "== $v"
  DSRT_LIB=$LIB timeout 300 python tests/gpu_pose_split.py 2>&1 | tee gpurun_out/pose_split_${TAG}_$v.txt
done
[ -n "$NO_NCU" ] && exit 0
CMD="python tests/gpu_pose_frame.py 98 64 10"
$CMD && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_pose98_$TAG.csv $CMD > gpurun_out/ncu_pose98_$TAG.log 2>&1
echo ncu_rc=$?
python profiles/summarize_ncu.py launches gpurun_out/launches_pose98_$TAG.csv | tee gpurun_out/launches_pose98_$TAG.txt
