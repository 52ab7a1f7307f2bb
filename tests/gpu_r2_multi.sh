#!/bin/bash
# Round-2 multi-GPU call (not a pytest file): N GPUs of one box.
#   gpurun --gpus N --timeout 1500 -- 'bash tests/gpu_r2_multi.sh N <tag>'
# 1. the two-GPU identity test of the parity suite (frame-parallel files / sample-sharded accumulation vs one GPU)
# 2. bench.py at N GPUs (frame-parallel C2 + the c4_sample sub-record with the timed NCCL reduce)
# 3. BASELINE config 3, the whole 99-pose sequence with PPM + PNG written: the CLI at 1 and N GPUs, and bench.py --config c3
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$N" -ge 2 ] && [ -z "$ONLY_C3" ]; then
  timeout 600 python -m pytest tests/test_gpu_pipeline_nets.py -m gpu -q -k two_gpus -s > gpurun_out/pytest_2gpu_$TAG.log 2>&1; echo pytest_2gpu_rc=$?
  tail -3 gpurun_out/pytest_2gpu_$TAG.log; cat gpurun_out/two_gpu_identity.log
fi
if [ -z "$ONLY_C3" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo bench_rc=$?
tail -1 gpurun_out/bench_n${N}_$TAG.json | cut -c1-600; tail -2 gpurun_out/bench_n${N}_$TAG.err
fi
RT=deep-space-ray-tracer_b200/ray_tracer
SEQ="--input_txt tests/golden/rendezvous_1s_dt0_01s.txt --width 1920 --height 1080 --spp 250 --max_depth 50 --synthetic_tris 1000000"
for G in 1 $N; do
  [ "$G" = 1 ] && [ "$N" != 1 ] && [ -n "$SKIP1" ] && continue
  T0=$(date +%s.%N)
  $RT $SEQ --output_dir /tmp/c3_out --gpus $G --stats_json gpurun_out/c3_cli_g${G}_$TAG.json > gpurun_out/c3_cli_g${G}_$TAG.log 2>&1; echo "cli_rc=$? gpus=$G wall=$(python -c "import sys,time; print(round(time.time() - float(sys.argv[1]), 2))" $T0) s"
  tail -1 gpurun_out/c3_cli_g${G}_$TAG.log; cat gpurun_out/c3_cli_g${G}_$TAG.json; ls /tmp/c3_out | wc -l
  [ "$G" = "$N" ] && break
done
rm -rf /tmp/c3_out
if [ "$N" = 1 ]; then
  timeout 600 python bench.py --config c3 > gpurun_out/bench_c3_n1_$TAG.json 2> gpurun_out/bench_c3_n1_$TAG.err; echo c3_rc=$?
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config c3 > gpurun_out/bench_c3_n${N}_$TAG.json 2> gpurun_out/bench_c3_n${N}_$TAG.err; echo c3_rc=$?
fi
tail -1 gpurun_out/bench_c3_n${N}_$TAG.json | cut -c1-900; tail -2 gpurun_out/bench_c3_n${N}_$TAG.err
