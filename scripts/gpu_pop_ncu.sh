#!/bin/bash
# usage: scripts/gpu_pop_ncu.sh <tag> [n_grbs] [kernel-regex] [count]
# population workload under ncu: launch list (durations) + one --set full capture of the matching kernels.
# Batches are held to ~10 GB of event buffers so that ncu's save / restore between replay passes stays cheap.
tag=${1:-rX}; n=${2:-256}; pat=${3:-"k_merge_deadtime|k_background|k_gap_prefix|k_trig_bin|k_trig_eval|k_sample_energy_env|k_source_thin"}; cnt=${4:-16}
mkdir -p gpurun_out
export COSMOGRB_B200_BATCH_BYTES=${COSMOGRB_B200_BATCH_BYTES:-10e9}
CMD="python scripts/pop_run.py $n"
$CMD > gpurun_out/pop_plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/pop_launches_${tag}.csv $CMD > gpurun_out/pop_ncu1_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$pat" -c $cnt -f -o gpurun_out/prof_pop_${tag} $CMD > gpurun_out/pop_ncu2_${tag}.log 2>&1
echo "pop ncu rc=$?"
tail -2 gpurun_out/pop_plain_${tag}.log
