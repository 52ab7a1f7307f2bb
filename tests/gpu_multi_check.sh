#!/bin/bash
# Multi-GPU + CLI check on a GPU box (run through gpurun --gpus 2).  Not a pytest file.
set -x
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_n${N}_frame.json 2> gpurun_out/bench_n${N}_frame.err; echo frame_rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --mode sample --width 3840 --height 2160 --spp 64 > gpurun_out/bench_n${N}_sample.json 2> gpurun_out/bench_n${N}_sample.err; echo sample_rc=$?
python bench.py --steps 3 --warmup 2 --width 3840 --height 2160 --spp 64 --no_cpu_baseline > gpurun_out/bench_n1_4k.json 2> gpurun_out/bench_n1_4k.err
cat gpurun_out/bench_n${N}_frame.json gpurun_out/bench_n${N}_sample.json gpurun_out/bench_n1_4k.json | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['scaling'], round(d['value']), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step e2e', d['e2e']['value'] and round(d['e2e']['value']), d['config']['parallelism'])
"
tail -3 gpurun_out/*.err
# CLI: frame-parallel over N GPUs, then sample-sharded with the NCCL reduce, compared with 1 GPU
RT=deep-space-ray-tracer_b200/ray_tracer
COMMON="--input_txt tests/golden/rendezvous_1s_dt0_01s.txt --width 400 --height 225 --spp 32 --max_depth 10 --synthetic_tris 100000 --frames 94:98 --no_png"
$RT $COMMON --output_dir gpurun_out/cli_1gpu --gpus 1 > gpurun_out/cli_1gpu.log 2>&1; echo cli1=$?
$RT $COMMON --output_dir gpurun_out/cli_frame --gpus $N > gpurun_out/cli_frame.log 2>&1; echo cliN=$?
$RT $COMMON --output_dir gpurun_out/cli_sample --gpus $N --shard sample > gpurun_out/cli_sample.log 2>&1; echo cliS=$?
for f in gpurun_out/cli_1gpu/*.ppm; do b=$(basename $f); cmp $f gpurun_out/cli_frame/$b && echo "frame-parallel $b identical"; python - <<PY
import numpy as np
a=np.fromfile("$f",np.uint8)[-400*225*3:].astype(int); b=np.fromfile("gpurun_out/cli_sample/$b",np.uint8)[-400*225*3:].astype(int)
print("sample-sharded $b: identical pixels", float((a==b).mean()), "max diff", int(abs(a-b).max()))
PY
done
tail -4 gpurun_out/cli_sample.log
# the reference's own main.cpp linked against libdsrt.so (no OBJ in the tree -> empty mesh, black frames)
(cd gpurun_out && mkdir -p dropin && cd dropin && timeout 120 ../../oracle/_ref/ray_tracer_dropin --output_dir out > dropin.log 2>&1; echo dropin=$?; tail -3 dropin.log; ls out | head -3)
(cd gpurun_out && mkdir -p binding && cd binding && timeout 120 ../../oracle/_ref/ray_tracer_binding --output_dir out > binding.log 2>&1; echo binding=$?; tail -3 binding.log; ls out | head -3)
rm -rf gpurun_out/cli_*/ gpurun_out/dropin/out gpurun_out/binding/out
