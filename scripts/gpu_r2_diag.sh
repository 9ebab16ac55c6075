#!/bin/bash
# usage: scripts/gpu_r2_diag.sh <tag> <kernel regex>: merge-kernel phase timing (MD_TIMING variant if built) + ncu --set full on C2 and a 128-burst population
tag=$1; pat=$2
mkdir -p gpurun_out
if [ -f cosmogrb_b200/_lib/variants/mdtiming.so ]; then
COSMOGRB_B200_LIB=$PWD/cosmogrb_b200/_lib/variants/mdtiming.so timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 100 > gpurun_out/bench_timing_${tag}.json 2> gpurun_out/bench_timing_${tag}.err
grep MDT gpurun_out/bench_timing_${tag}.err | tail -30
fi
scripts/gpu_r2_ncu_both.sh ${tag} "$pat"
