#!/bin/bash
# Quick GPU check of a kernel change: parity suite + the bench line (+ optional variant libs).
#   gpurun --timeout 900 -- 'bash tests/gpu_quick.sh <tag> [variant.so ...]'
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_rc=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --no_cpu_baseline --no_ref_gpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("$TAG", round(d["value"]), "Mrays/s", round(d["ms_per_step"],2), "ms  trav", round(r["kernel_ms_per_step"],2), "primary", round(r["k_primary_ms_per_step"],2), "shade", round(r["k_shade_ms_per_step"],2), "e2e", round(d["e2e"]["value"]))
PY
for lib in "$@"; do
  DSRT_LIB=$lib timeout 300 python bench.py --no_cpu_baseline --no_ref_gpu --no_e2e > gpurun_out/bench_${TAG}_$(basename $lib .so).json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_$(basename $lib .so).json").read().strip().splitlines()[-1])
r=d["roofline"]; print("$lib", round(d["value"]), "Mrays/s", round(d["ms_per_step"],2), "ms  trav", round(r["kernel_ms_per_step"],2), "primary", round(r["k_primary_ms_per_step"],2), "shade", round(r["k_shade_ms_per_step"],2))
PY
done
