#!/bin/bash
mkdir -p gpurun_out
COSMOGRB_B200_LIB=$PWD/cosmogrb_b200/_lib/variants/mdtiming.so timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 100 > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err
grep MDT gpurun_out/bench_timing.err | tail -24
