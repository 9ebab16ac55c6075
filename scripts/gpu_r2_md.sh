#!/bin/bash
# usage: scripts/gpu_r2_md.sh <tag> <pytest -k expr> <variant...>: tests, bench base + variants (C2 + 300-burst population), phase timing
tag=$1; kexpr=$2; shift; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -12 gpurun_out/pytest_${tag}.log
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --population 300 > gpurun_out/bench_${tag}_$1.json 2> gpurun_out/bench_${tag}_$1.err; echo "$1 rc=$?"; python scripts/show_bench.py gpurun_out/bench_${tag}_$1.json | grep -E "value|k_merge|k_source_thin|k_thin_nodes|k_sample_energy_env|k_background|k_gap_prefix|k_trig|population"; }
run base
for v in "$@"; do COSMOGRB_B200_LIB=$PWD/cosmogrb_b200/_lib/variants/$v.so run $v; done
if [ -f cosmogrb_b200/_lib/variants/mdtiming.so ]; then
COSMOGRB_B200_LIB=$PWD/cosmogrb_b200/_lib/variants/mdtiming.so timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 100 > gpurun_out/bench_timing_${tag}.json 2> gpurun_out/bench_timing_${tag}.err
grep MDT gpurun_out/bench_timing_${tag}.err | tail -24
fi
