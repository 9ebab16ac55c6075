#!/bin/bash
# usage: scripts/gpu_variants2.sh <pytest -k expr or ""> <variant...> : GPU tests on the base library, then the per-kernel population timings
# (3000 bursts) and the C2 step for the base library and each variant under cosmogrb_b200/_lib/variants/
kexpr=$1; shift
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then timeout 900 python -m pytest tests -m gpu -q -x -k "$kexpr" 2>&1 | tail -4; fi
run() { echo "== $1"; timeout 300 python scripts/pop_diag.py 3000 16e9 2>/dev/null | tail -2 | cut -c1-420
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --population-total 0 2>/dev/null > gpurun_out/bench_var_$1.json
  python scripts/show_bench.py gpurun_out/bench_var_$1.json | grep -E "value|k_merge_deadtime|k_source_thin|k_sample_energy_env|k_gap_prefix|k_source_gather" | cut -c1-160; }
run base
for v in "$@"; do COSMOGRB_B200_LIB=$PWD/cosmogrb_b200/_lib/variants/$v.so run $v; done
