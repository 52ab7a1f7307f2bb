#!/bin/bash
# CUDA start-up on a multi-GPU box (not a pytest file): where the seconds before the first frame go.
#   gpurun --gpus 8 --timeout 300 -- 'bash tests/gpu_startup.sh'
mkdir -p gpurun_out
{
nvidia-smi -q | grep -i -m2 "persistence"
N=$(nvidia-smi -L | wc -l)
echo "== one process, every device visible, contexts one after another"
tests/c/cuda_startup seq
echo "== one process, every device visible, one thread per device"
tests/c/cuda_startup par
echo "== one process, one device visible"
CUDA_VISIBLE_DEVICES=0 tests/c/cuda_startup one
echo "== $N processes at once, each with one device visible"
T0=$(date +%s.%N)
for g in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$g tests/c/cuda_startup one $T0 > /tmp/start_$g.txt & done
wait
cat /tmp/start_*.txt | grep "^one (CUDA"
echo "== $N processes, staggered by 0.3 s"
T0=$(date +%s.%N)
for g in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$g tests/c/cuda_startup one $T0 > /tmp/start_$g.txt & sleep 0.3; done
wait
cat /tmp/start_*.txt | grep "^one (CUDA"
} 2>&1 | tee gpurun_out/cuda_startup.txt
