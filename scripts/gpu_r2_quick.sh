#!/bin/bash
# usage: scripts/gpu_r2_quick.sh <tag> [pytest -k expr] [ncu kernel regex]
tag=${1:-R2x}; kexpr=$2; pat=$3
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
else
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
fi
tail -12 gpurun_out/pytest_${tag}.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_${tag}.err
python scripts/show_bench.py gpurun_out/bench_${tag}.json
if [ -n "$pat" ]; then
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 1 -c 2 -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu2_${tag}.log 2>&1
echo "ncu rc=$?"
fi
