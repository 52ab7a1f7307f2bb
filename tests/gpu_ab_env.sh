#!/bin/bash
# A/B of run-time switches on a GPU box: parity suite once, then the bench line per environment setting.
#   gpurun --timeout 900 -- 'bash tests/gpu_ab_env.sh <tag> "VAR=a" "VAR=b" ...'
TAG=${1:-ab}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_rc=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
k=0
for setting in "$@"; do
  k=$((k+1))
  env $setting timeout 300 python bench.py --no_cpu_baseline --no_ref_gpu $BENCH_ARGS > gpurun_out/bench_${TAG}_$k.json 2> gpurun_out/bench_${TAG}_$k.err; echo "bench_rc=$? ($setting)"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_$k.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("$setting", round(d["value"]), "Mrays/s", round(d["ms_per_step"],2), "ms  trav", round(r["kernel_ms_per_step"],2), "primary", round(r["k_primary_ms_per_step"],2), "shade", round(r["k_shade_ms_per_step"],2), "bvh", d["config"]["bvh"])
PY
done
