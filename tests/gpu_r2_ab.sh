#!/bin/bash
# A/B of kernel variants (deep-space-ray-tracer_b200/build/variants/libdsrt_<tag>.so) on a GPU box; not a pytest file.
#   gpurun --timeout 1800 -- 'bash tests/gpu_r2_ab.sh <tag> "<variant tags>" [c5]'
TAG=${1:-ab}; VARS=${2:-}; C5=${3:-}
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_rc=$?
tail -4 gpurun_out/pytest_gpu_$TAG.log
for v in main $VARS; do
  LIB=""; [ "$v" != main ] && LIB=$PWD/deep-space-ray-tracer_b200/build/variants/libdsrt_$v.so
  DSRT_LIB=$LIB timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_ref_gpu --no_sub --no_e2e > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err; echo "bench_rc=$? ($v)"
  [ -n "$C5" ] && { DSRT_LIB=$LIB timeout 300 python bench.py --config c5 --steps 3 --warmup 2 --no_cpu_baseline --no_ref_gpu --no_sub --no_e2e > gpurun_out/bench_${TAG}_${v}_c5.json 2> gpurun_out/bench_${TAG}_${v}_c5.err; echo "c5_rc=$? ($v)"; }
  python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_$v.json", "gpurun_out/bench_${TAG}_${v}_c5.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        continue
    r=d["roofline"]; print("$v", f.endswith("c5.json") and "C5" or "C2", round(d["value"]), "Mrays/s", round(d["ms_per_step"],2), "ms  trav", round(r["kernel_ms_per_step"],2), "primary", round(r["k_primary_ms_per_step"],2), "shade", round(r["k_shade_ms_per_step"],2))
PY
done
