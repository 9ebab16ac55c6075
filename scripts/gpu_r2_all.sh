#!/bin/bash
# usage: scripts/gpu_r2_all.sh <tag> <variant...>: the whole GPU test suite, then the bench (C2 + 300-burst population) for base + variants
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -15 gpurun_out/pytest_${tag}.log
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --population 300 > gpurun_out/bench_${tag}_$1.json 2> gpurun_out/bench_${tag}_$1.err; echo "$1 rc=$?"; tail -3 gpurun_out/bench_${tag}_$1.err | cut -c1-300; python scripts/show_bench.py gpurun_out/bench_${tag}_$1.json | grep -E "value|k_merge|k_source_thin|k_thin_nodes|k_sample_energy_env|k_background|k_gap_prefix|k_trig|population"; }
run base
for v in "$@"; do COSMOGRB_B200_LIB=$PWD/cosmogrb_b200/_lib/variants/$v.so run $v; done
