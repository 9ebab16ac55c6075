#!/bin/bash
# usage: scripts/gpu_profile.sh <tag>   (run on the GPU box through gpurun)
# 1. GPU test suite, 2. full bench line, 3. ncu launch list + one --set full capture of one warm step.
tag=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 0"
$CMD > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 17 --csv \
    --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu1_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 18 -c 17 -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu2_${tag}.log 2>&1
echo "profile rc=$?"
tail -3 gpurun_out/pytest_${tag}.log
