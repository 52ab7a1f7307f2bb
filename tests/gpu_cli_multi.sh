#!/bin/bash
# CLI on N GPUs of one box (not a pytest file): frame-parallel and sample-sharded runs against the
# 1-GPU frames, with and without per-worker GPU isolation.
#   gpurun --gpus N --timeout 600 -- 'bash tests/gpu_cli_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
RT=deep-space-ray-tracer_b200/ray_tracer
COMMON="--input_txt tests/golden/rendezvous_1s_dt0_01s.txt --width 400 --height 225 --spp 32 --max_depth 10 --synthetic_tris 100000 --frames 90:98 --no_png"
$RT $COMMON --output_dir gpurun_out/cli_1gpu --gpus 1 > gpurun_out/cli_1gpu.log 2>&1; echo cli1=$?
for iso in 1 0; do
  DSRT_CLI_ISOLATE=$iso $RT $COMMON --output_dir gpurun_out/cli_frame$iso --gpus $N --stats_json gpurun_out/cli_frame$iso.json > gpurun_out/cli_frame$iso.log 2>&1; echo "frame iso=$iso rc=$?"
  DSRT_CLI_ISOLATE=$iso $RT $COMMON --output_dir gpurun_out/cli_sample$iso --gpus $N --shard sample --stats_json gpurun_out/cli_sample$iso.json > gpurun_out/cli_sample$iso.log 2>&1; echo "sample iso=$iso rc=$?"
  cat gpurun_out/cli_frame$iso.json gpurun_out/cli_sample$iso.json
  same=0; total=0
  for f in gpurun_out/cli_1gpu/*.ppm; do b=$(basename $f); total=$((total+1)); cmp -s $f gpurun_out/cli_frame$iso/$b && same=$((same+1)); done
  echo "iso=$iso frame-parallel: $same of $total frames byte-identical to the 1-GPU run"
  python - <<PY
import numpy as np, glob, os
worst = 0; n = 0
for f in sorted(glob.glob("gpurun_out/cli_1gpu/*.ppm")):
    g = "gpurun_out/cli_sample$iso/" + os.path.basename(f)
    a = np.fromfile(f, np.uint8)[-400*225*3:].astype(int); b = np.fromfile(g, np.uint8)[-400*225*3:].astype(int)
    worst = max(worst, int(abs(a - b).max())); n += 1
print("iso=$iso sample-sharded:", n, "frames, max byte difference to the 1-GPU run", worst)
PY
done
tail -3 gpurun_out/cli_sample1.log
rm -rf gpurun_out/cli_1gpu gpurun_out/cli_frame? gpurun_out/cli_sample?
