#!/bin/bash
tag=${1:-R2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -8 gpurun_out/pytest_${tag}.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
scripts/gpu_pop_ncu.sh ${tag} 256
