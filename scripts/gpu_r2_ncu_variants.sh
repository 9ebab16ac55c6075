#!/bin/bash
# usage: scripts/gpu_r2_ncu_variants.sh <tag> <kernel regex> <variant...>: bench for base + variants, then ncu --set full of base
tag=$1; pat=$2; shift; shift
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${tag}_$1.json 2> gpurun_out/bench_${tag}_$1.err; echo "$1 rc=$?"; python scripts/show_bench.py gpurun_out/bench_${tag}_$1.json | grep -E "value|k_merge_deadtime|k_source_thin|k_thin_nodes|k_sample_energy_env|k_background|k_gap_prefix|k_trig|population"; }
run base
for v in "$@"; do COSMOGRB_B200_LIB=$PWD/cosmogrb_b200/_lib/variants/$v.so run $v; done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 1 -c 2 -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu2_${tag}.log 2>&1
echo "ncu rc=$?"
