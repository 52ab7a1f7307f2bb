#!/bin/bash
# ncu launch list of the bench command itself at its full sample count (GPU box; not a pytest file).
#   gpurun --timeout 600 -- 'bash tests/gpu_r2_launchlist.sh <tag>'
TAG=${1:-ll}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no_cpu_baseline --no_e2e --no_ref_gpu --no_sub"
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err && cut -c1-200 gpurun_out/plain_$TAG.json &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file /tmp/launches_$TAG.csv $CMD > gpurun_out/ncu_ll_$TAG.log 2>&1
echo ncu_rc=$?
python profiles/summarize_ncu.py launches /tmp/launches_$TAG.csv > gpurun_out/launches_$TAG.txt; echo sum_rc=$?
grep -B1 -A12 "render-step kernels only" gpurun_out/launches_$TAG.txt
