#!/bin/bash
# usage: scripts/gpu_variants3.sh <variant...> : the C2 step (per-kernel live timings) and the per-kernel population timings
# (2000 bursts) for the base library and each variant under cosmogrb_b200/_lib/variants/ -- no tests, for tuning calls
mkdir -p gpurun_out
run() { echo "== $1"
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --population-total 0 2>/dev/null > gpurun_out/bench_var_$1.json
  python scripts/show_bench.py gpurun_out/bench_var_$1.json | grep -E "value|k_merge_deadtime|k_source_thin|k_sample_energy_env|k_gap_prefix|k_source_gather" | cut -c1-160
  timeout 200 python scripts/pop_diag.py 2000 16e9 2>/dev/null | tail -2 | cut -c1-420; }
run base
for v in "$@"; do COSMOGRB_B200_LIB=$PWD/cosmogrb_b200/_lib/variants/$v.so run $v; done
