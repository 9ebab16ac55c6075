#!/bin/bash
# usage: scripts/gpu_r2_ncu_both.sh <tag> <kernel regex>: ncu --set full of the matching kernels on one C2 step and on a 128-burst population
tag=$1; pat=$2
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 1 -c 1 -f -o gpurun_out/prof_${tag}_c2 $CMD > gpurun_out/ncu2_${tag}_c2.log 2>&1
echo "ncu c2 rc=$?"
export COSMOGRB_B200_BATCH_BYTES=${COSMOGRB_B200_BATCH_BYTES:-10e9}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$pat" -c 2 -f -o gpurun_out/prof_${tag}_pop python scripts/pop_run.py 128 > gpurun_out/ncu2_${tag}_pop.log 2>&1
echo "ncu pop rc=$?"
