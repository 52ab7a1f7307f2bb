#!/bin/bash
# ncu capture of the traversal kernels on BASELINE config 5 (10M triangles, the HBM-bound case).
#   gpurun --timeout 900 -- 'bash tests/gpu_profile_c5.sh <tag>'
TAG=${1:-c5}
mkdir -p gpurun_out
PROF="python bench.py --config c5 --steps 1 --warmup 1 --spp 16 --no_cpu_baseline --no_e2e --no_ref_gpu"
$PROF > gpurun_out/plain_c5_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_primary|k_trace|k_shade' -c 5 -f -o gpurun_out/prof_c5_$TAG $PROF > gpurun_out/ncu_c5_$TAG.log 2>&1
echo ncu_rc=$?
tail -2 gpurun_out/ncu_c5_$TAG.log
