#!/bin/bash
# usage: scripts/gpu_ncu.sh <tag> <kernel-regex> [skip] [count]   one --set full capture of matching kernels of one bench step
tag=${1:-rX}; pat=$2; skip=${3:-1}; cnt=${4:-1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 0"
$CMD > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$pat" -s $skip -c $cnt -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu2_${tag}.log 2>&1
echo "ncu rc=$?"
