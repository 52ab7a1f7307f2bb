#!/bin/bash
# Round-2 final check on one GPU, in the order the driver runs things (not a pytest file).
#   gpurun --timeout 1500 -- 'bash tests/gpu_r2_final.sh <tag>'
TAG=${1:-fin}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_rc=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo smoke_rc=$?; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo ref_rc=$?; cut -c1-400 gpurun_out/bench_ref_$TAG.json
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?; cut -c1-300 gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
