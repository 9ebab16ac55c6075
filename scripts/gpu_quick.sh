#!/bin/bash
# usage: scripts/gpu_quick.sh <tag> [kernel-regex]   GPU tests + bench (no cpu leg) + optional ncu capture of matching kernels
tag=${1:-rX}; pat=$2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -15 gpurun_out/pytest_${tag}.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${tag}.err
if [ -n "$pat" ]; then
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 0"
$CMD > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 3 -c 3 -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu2_${tag}.log 2>&1
echo "ncu rc=$?"
fi
