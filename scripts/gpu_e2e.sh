#!/bin/bash
# e2e arm only (C2), for a few pipeline settings
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --population-total 0 2>/dev/null > gpurun_out/bench_e2e.json
  python - <<'P'
import json
b=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
e=b['e2e']; print(f"   e2e {e['value']:.4g} ev/s  d2h {e['d2h_gbs_per_gpu']:.1f} GB/s of {e['d2h_yardstick_gbs_per_gpu']:.1f} = {e['frac_of_d2h_yardstick']:.3f}; value {b['value']:.4g}")
P
}
run CGRB_PIPELINE_CHUNKS=7
run CGRB_PIPELINE_CHUNKS=14
run CGRB_PIPELINE_CHUNKS=4
run CGRB_PIPELINE_CHUNKS=7
