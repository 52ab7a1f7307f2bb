#!/bin/bash
# The other BASELINE configurations on one GPU (not a pytest file): C1 (the CPU-runnable case) and C4 (4K, 1000 spp).
TAG=${1:-cfg}
mkdir -p gpurun_out
timeout 300 python bench.py --tris 10000 --width 400 --height 225 --spp 16 --depth 10 --steps 20 --warmup 3 --no_sub > gpurun_out/bench_c1_$TAG.json 2> gpurun_out/bench_c1_$TAG.err; echo c1_rc=$?
timeout 300 python bench.py --config c4 --steps 3 --warmup 1 --no_sub --no_cpu_baseline --no_ref_gpu > gpurun_out/bench_c4_n1_$TAG.json 2> gpurun_out/bench_c4_n1_$TAG.err; echo c4_rc=$?
python - <<PY
import json
for f in ("gpurun_out/bench_c1_$TAG.json", "gpurun_out/bench_c4_n1_$TAG.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), "Mrays/s", round(d["ms_per_step"], 3), "ms; e2e", d["e2e"]["ms_per_step"], "cpu", (d.get("cpu_baseline") or {}).get("value"), "refgpu", (d.get("ref_gpu_kernel") or {}).get("kernel_ms_per_frame"))
PY
