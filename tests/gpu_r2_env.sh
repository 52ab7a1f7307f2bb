#!/bin/bash
# A/B of run-time switches (environment settings) on a GPU box; not a pytest file.
#   gpurun --timeout 1500 -- 'bash tests/gpu_r2_env.sh <tag> [c5] -- "VAR=a" "VAR=b" ...'
TAG=${1:-env}; shift; C5=""; [ "$1" = c5 ] && { C5=1; shift; }; [ "$1" = "--" ] && shift
mkdir -p gpurun_out
k=0
for setting in "$@"; do
  k=$((k+1))
  env $setting timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_ref_gpu --no_sub --no_e2e > gpurun_out/bench_${TAG}_$k.json 2> gpurun_out/bench_${TAG}_$k.err; echo "bench_rc=$? ($setting)"
  [ -n "$C5" ] && { env $setting timeout 300 python bench.py --config c5 --steps 3 --warmup 2 --no_cpu_baseline --no_ref_gpu --no_sub --no_e2e > gpurun_out/bench_${TAG}_${k}_c5.json 2> gpurun_out/bench_${TAG}_${k}_c5.err; echo "c5_rc=$?"; }
  python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_$k.json", "gpurun_out/bench_${TAG}_${k}_c5.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        continue
    r=d["roofline"]; print("$setting", f.endswith("c5.json") and "C5" or "C2", round(d["value"]), "Mrays/s", round(d["ms_per_step"],2), "ms  trav", round(r["kernel_ms_per_step"],2), "primary", round(r["k_primary_ms_per_step"],2), "shade", round(r["k_shade_ms_per_step"],2), "nodes", d["config"]["bvh"]["wide_nodes"], "sah", round(d["config"]["bvh"]["sah"],2))
PY
done
