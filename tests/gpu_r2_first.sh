#!/bin/bash
# Round-2 GPU check (not a pytest file): the parity suite, then short bench lines of the configurations
# the round tracks.    gpurun --timeout 1500 -- 'bash tests/gpu_r2_first.sh <tag>'
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_rc=$?
tail -25 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_ref_gpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --config c5 --steps 5 --warmup 3 --no_cpu_baseline --no_ref_gpu --no_e2e > gpurun_out/bench_c5_$TAG.json 2> gpurun_out/bench_c5_$TAG.err; echo bench_c5_rc=$?
cat gpurun_out/bench_c5_$TAG.json; tail -3 gpurun_out/bench_c5_$TAG.err
