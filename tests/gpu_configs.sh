#!/bin/bash
# BASELINE.json configurations on a GPU box (not a pytest file).
#   gpurun --timeout 1500 -- 'bash tests/gpu_configs.sh <tag> 1'          single GPU: C2, C5, C4 (1 GPU), C3 subset
#   gpurun --gpus N --timeout 1500 -- 'bash tests/gpu_configs.sh <tag> N'  N GPUs: C2 frame-parallel, C4 sample-sharded, C3 CLI
TAG=${1:-cfg}; N=${2:-1}
mkdir -p gpurun_out
RT=deep-space-ray-tracer_b200/ray_tracer
POSES=tests/golden/rendezvous_1s_dt0_01s.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() { python - "$@" <<'PY'
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable", e); continue
    if "value" in d:
        r = d.get("roofline", {})
        print(p.split("/")[-1], d["n_gpus"], "GPU", d["scaling"], round(d["value"]), d["unit"], round(d["ms_per_step"], 2), "ms/step",
              "frac", round(r.get("frac", 0), 3), "e2e", d["e2e"]["value"] and round(d["e2e"]["value"]),
              "ref_gpu", d.get("ref_gpu_kernel", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"), d["config"].get("frames_per_s"))
    else:
        print(p.split("/")[-1], d)
PY
}
if [ "$N" = "1" ]; then
  python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo c2_rc=$?
  python bench.py --config c5 --steps 5 --no_cpu_baseline > gpurun_out/bench_${TAG}_c5.json 2> gpurun_out/bench_${TAG}_c5.err; echo c5_rc=$?
  python bench.py --config c4 --steps 2 --warmup 1 --no_cpu_baseline --no_ref_gpu > gpurun_out/bench_${TAG}_c4_1gpu.json 2> gpurun_out/bench_${TAG}_c4_1gpu.err; echo c4_rc=$?
  show gpurun_out/bench_${TAG}_c2.json gpurun_out/bench_${TAG}_c5.json gpurun_out/bench_${TAG}_c4_1gpu.json
  # C3: the pose sequence through the CLI at 1080p, 250 spp, the reference's default depth 50 (first 20 frames on one GPU)
  $RT --input_txt $POSES --output_dir gpurun_out/c3_out --width 1920 --height 1080 --spp 250 --frames 0:20 --gpus 1 --stats_json gpurun_out/c3_${TAG}_1gpu.json > gpurun_out/c3_${TAG}_1gpu.log 2>&1; echo c3_rc=$?
  cat gpurun_out/c3_${TAG}_1gpu.json; ls gpurun_out/c3_out | head -3; cp gpurun_out/c3_out/frame_0019.png gpurun_out/c3_frame_0019.png; rm -rf gpurun_out/c3_out
else
  $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_c2_n$N.json 2> gpurun_out/bench_${TAG}_c2_n$N.err; echo c2n_rc=$?
  $TR --master-port 29512 bench.py --gpus $N --config c4 --mode sample --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_c4_n$N.json 2> gpurun_out/bench_${TAG}_c4_n$N.err; echo c4n_rc=$?
  show gpurun_out/bench_${TAG}_c2_n$N.json gpurun_out/bench_${TAG}_c4_n$N.json
  tail -2 gpurun_out/bench_${TAG}_c2_n$N.err gpurun_out/bench_${TAG}_c4_n$N.err
  # C3: all 99 poses, frame-parallel over the N GPUs
  $RT --input_txt $POSES --output_dir gpurun_out/c3_out --width 1920 --height 1080 --spp 250 --gpus $N --stats_json gpurun_out/c3_${TAG}_n$N.json > gpurun_out/c3_${TAG}_n$N.log 2>&1; echo c3n_rc=$?
  cat gpurun_out/c3_${TAG}_n$N.json; ls gpurun_out/c3_out | wc -l; rm -rf gpurun_out/c3_out
  # C4 through the CLI: one 4K frame, 1000 spp, sample-sharded with the NCCL reduce
  $RT --input_txt $POSES --output_dir gpurun_out/c4_out --width 3840 --height 2160 --spp 1000 --frames 0:1 --gpus $N --shard sample --stats_json gpurun_out/c4cli_${TAG}_n$N.json > gpurun_out/c4cli_${TAG}_n$N.log 2>&1; echo c4cli_rc=$?
  cat gpurun_out/c4cli_${TAG}_n$N.json; rm -rf gpurun_out/c4_out
fi
