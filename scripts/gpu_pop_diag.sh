#!/bin/bash
mkdir -p gpurun_out
POP_DIAG_BATCHES=1 timeout 300 python scripts/pop_diag.py ${1:-3000} ${2:-32e9} 2>/dev/null | tail -40
