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
# DRAM traffic of the traversal launches of one step of the DEFAULT workload and of C5 (bench.py's roofline.traffic).
# Only the DRAM / L2 byte counters and the duration are collected (the figures `--set full` would give, in 2 replay passes
# instead of ~40: a C5 step touches 19 GB that ncu saves and restores around every pass).  The reports are summarised HERE:
# gpurun brings back at most 64 MiB.
WL2="pose 0 of rendezvous_1s_dt0_01s.txt, synthetic 1000000-triangle station, 1920x1080, 64 spp, depth 10, directional sun + shadow rays"
WL5="pose 0 of rendezvous_1s_dt0_01s.txt, synthetic 10000000-triangle station (diffuse only), 1920x1080, 256 spp, depth 16, directional sun + shadow rays"
if [ -z "$NO_TRAFFIC" ]; then
TR="python bench.py --steps 1 --warmup 1 --no_cpu_baseline --no_e2e --no_ref_gpu --no_sub"
$TR > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k 'regex:k_primary|k_trace|k_tail' -c 80 -f -o /tmp/traffic_$TAG $TR > gpurun_out/ncu3_$TAG.log 2>&1
echo ncu3_rc=$?
python profiles/summarize_ncu.py traffic /tmp/traffic_$TAG.ncu-rep "$WL2" 1 > gpurun_out/trace_ncu_summary_$TAG.json; echo sum3_rc=$?
# C5: only the second (timed) step's two batches; the first 1 + 1 steps x 2 batches are skipped by launch count
TR5="python bench.py --config c5 --steps 1 --warmup 0 --no_cpu_baseline --no_e2e --no_ref_gpu --no_sub"
$TR5 > gpurun_out/plain5_$TAG.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k 'regex:k_primary|k_trace|k_tail' -c 100 -f -o /tmp/traffic_c5_$TAG $TR5 > gpurun_out/ncu4_$TAG.log 2>&1
echo ncu4_rc=$?
python profiles/summarize_ncu.py traffic /tmp/traffic_c5_$TAG.ncu-rep "$WL5" 2 > gpurun_out/trace_ncu_summary_c5_$TAG.json; echo sum4_rc=$?
fi
du -sh gpurun_out
