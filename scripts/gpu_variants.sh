#!/bin/bash
# usage: scripts/gpu_variants.sh <tag> <variant...>   bench line (no cpu / e2e legs) for the default build and each
# variant under cosmogrb_b200/_lib/variants/; "batch=<bytes>" runs the default build with that batch size
tag=$1; shift
mkdir -p gpurun_out
run() { timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${tag}_$1.json 2> gpurun_out/bench_${tag}_$1.err; echo "$1 rc=$?"; }
run base
for v in "$@"; do
  case $v in
    batch=*) COSMOGRB_B200_BATCH_BYTES=${v#batch=} run $v ;;
    *) COSMOGRB_B200_LIB=$PWD/cosmogrb_b200/_lib/variants/$v.so run $v ;;
  esac
done
