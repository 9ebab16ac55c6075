#!/bin/bash
# usage: scripts/gpu_check.sh <tag> [pytest -k expression] : selected GPU tests, the per-batch population timings, a short bench
tag=${1:-chk}; kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then timeout 900 python -m pytest tests -m gpu -q -x -k "$kexpr" 2>&1 | tail -5; else timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5; fi
POP_DIAG_BATCHES=1 timeout 300 python scripts/pop_diag.py 3000 16e9 2>/dev/null > gpurun_out/pop_diag_${tag}.log; head -2 gpurun_out/pop_diag_${tag}.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --population-total ${POPTOTAL:-20000} > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_${tag}.json | head -60
