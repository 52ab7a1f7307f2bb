#!/bin/bash
# Round-2 GPU call (not a pytest file): parity suite, the bench line, then the ncu launch list and a full capture of
# the traversal / shade kernels of a shortened bench command.
#   gpurun --timeout 1800 -- 'bash tests/gpu_r2_profile.sh <tag>'
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_rc=$?
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
PROF="python bench.py --steps 1 --warmup 1 --spp 16 --no_cpu_baseline --no_e2e --no_ref_gpu --no_sub"
$PROF > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $PROF > gpurun_out/ncu1_$TAG.log 2>&1
echo ncu1_rc=$?
$PROF > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_primary|k_trace|k_shade' -c 7 -f -o gpurun_out/prof_$TAG $PROF > gpurun_out/ncu2_$TAG.log 2>&1
echo ncu2_rc=$?
tail -2 gpurun_out/ncu2_$TAG.log
# DRAM traffic of the traversal launches of one step of the DEFAULT workload and of C5 (bench.py's roofline.traffic)
TR="python bench.py --steps 1 --warmup 1 --no_cpu_baseline --no_e2e --no_ref_gpu --no_sub"
$TR > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:k_primary|k_trace|k_tail' -c 80 -f -o gpurun_out/traffic_$TAG $TR > gpurun_out/ncu3_$TAG.log 2>&1
echo ncu3_rc=$?
TR5="python bench.py --config c5 --steps 1 --warmup 1 --no_cpu_baseline --no_e2e --no_ref_gpu --no_sub"
$TR5 > gpurun_out/plain5_$TAG.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:k_primary|k_trace|k_tail' -c 200 -f -o gpurun_out/traffic_c5_$TAG $TR5 > gpurun_out/ncu4_$TAG.log 2>&1
echo ncu4_rc=$?
