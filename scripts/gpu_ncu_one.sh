#!/bin/bash
# usage: scripts/gpu_ncu_one.sh <tag> <kernel regex> [n_bursts] [count]: ncu --set full of matching kernels on a small population; csv pages only
tag=$1; pat=$2; n=${3:-128}; cnt=${4:-2}
mkdir -p gpurun_out
export COSMOGRB_B200_BATCH_BYTES=${COSMOGRB_B200_BATCH_BYTES:-10e9}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -c $cnt -f -o /tmp/prof_${tag} python scripts/pop_run.py $n > gpurun_out/ncu_${tag}.log 2>&1
echo "ncu rc=$?"
python scripts/ncu_summary.py /tmp/prof_${tag}.ncu-rep > gpurun_out/one_${tag}_summary.txt
python scripts/ncu_hot.py /tmp/prof_${tag}.ncu-rep "$pat" 70 > gpurun_out/one_${tag}_hot.txt
head -30 gpurun_out/one_${tag}_summary.txt
