#!/bin/bash
# population wall time with one and two compute streams (same box, back to back)
for s in 1 2 1 2; do echo "== streams $s"; COSMOGRB_B200_STREAMS=$s timeout 300 python scripts/pop_diag.py ${1:-20000} 32e9 2>/dev/null | head -1 | cut -c1-200; done
timeout 600 python -m pytest tests -m gpu -q -x -k "universe or trigger" 2>&1 | tail -2
