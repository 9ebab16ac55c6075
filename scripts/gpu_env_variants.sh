#!/bin/bash
# energy kernel: resident CTAs per SM and photons per work item (C2 step, no population)
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --population-total 0 2>/dev/null > gpurun_out/bench_env.json
  python scripts/show_bench.py gpurun_out/bench_env.json | grep -E "value|k_sample_energy_env " | cut -c1-160; }
run CGRB_ENV_MIN_CTAS=3
run CGRB_ENV_MIN_CTAS=4
run CGRB_ENV_MIN_CTAS=2
run CGRB_ENERGY_CHUNK=8192
run CGRB_ENERGY_CHUNK=32768
