#!/bin/bash
# usage: scripts/build_variant.sh <name> <nvcc -D flags...>   -> cosmogrb_b200/_lib/variants/<name>.so
# (run a tuning variant with COSMOGRB_B200_LIB=cosmogrb_b200/_lib/variants/<name>.so python bench.py ...)
name=$1; shift
mkdir -p cosmogrb_b200/_lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
    -I include -o cosmogrb_b200/_lib/variants/$name.so cosmogrb_b200/csrc/cgrb_kernels.cu
