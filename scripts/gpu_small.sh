#!/bin/bash
# small-batch latency (C1, C3) + population (default settings)
run() { echo "== $*"; env "$@" timeout 300 python scripts/config_bench.py 2>/dev/null > gpurun_out/configs_cur.json; python -c "
import json,sys
d=json.load(open('gpurun_out/configs_cur.json')); c=d['C1']; print('   C1 %.4f ms per burst, C3 %.4f ms;'%(c['gpu_ms_per_burst'], d['C3']['gpu_ms']), sorted(c['gpu_kernels_ms'].items(), key=lambda kv:-kv[1])[:6])"
  env "$@" timeout 300 python scripts/pop_diag.py 10000 32e9 2>/dev/null | head -1 | cut -c1-130; }
run CGRB_DUMMY=1
run CGRB_DUMMY=1
