#!/bin/bash
# usage: scripts/gpu_round2.sh <tag> [pop_bursts]   (on the GPU box through gpurun)
# 1. GPU test suite  2. the full bench line (C2 + 100k-burst population + cpu leg)
# 3. C2: ncu launch list of the whole command + --set full of one warm step (own kernels only, "k_*")
# 4. population: launch list + --set full of the hot kernels on a <pop_bursts>-burst population
tag=${1:-R2}; npop=${2:-256}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -6 gpurun_out/pytest_${tag}.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_${tag}.err | cut -c1-300
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --population-total 0"
timeout 300 $CMD > gpurun_out/plain_${tag}.json 2> gpurun_out/plain_${tag}.err || { echo "plain failed"; exit 1; }
N=$(python -c "import json;print(json.load(open('gpurun_out/plain_${tag}.json'))['gpu_launches'])")
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv \
    --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu1_${tag}.log 2>&1
echo "launch list rc=$? (N=$N per step)"
TOT=$(grep -c '"gpu__time_duration.sum"' gpurun_out/launches_${tag}.csv)
SKIP=$((TOT - N))
echo "own launches in the command: $TOT, skipping $SKIP"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s $SKIP -c $N -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu2_${tag}.log 2>&1
echo "full capture rc=$?"
# gpurun brings back at most 64 MiB: keep the csv pages, and the report itself only while it is small
post() { rep=$1; stem=$2; shift 2
  ncu -i $rep --page raw --csv > gpurun_out/${stem}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py $rep > gpurun_out/${stem}_summary.txt 2>/dev/null
  for k in "$@"; do python scripts/ncu_hot.py $rep "$k" 60 > gpurun_out/${stem}_hot_${k//[^a-z_]/}.txt 2>/dev/null; done
  sz=$(stat -c %s $rep); if [ "$sz" -gt 20000000 ]; then rm -f $rep; echo "$rep dropped ($sz bytes)"; fi; }
post gpurun_out/prof_${tag}.ncu-rep c2_${tag} k_sample_energy_env k_source_thin k_merge_deadtime k_gap_prefix
if [ "$npop" -gt 0 ]; then
export COSMOGRB_B200_BATCH_BYTES=${COSMOGRB_B200_BATCH_BYTES:-10e9}
PCMD="python scripts/pop_run.py $npop"
timeout 300 $PCMD > gpurun_out/pop_plain_${tag}.log 2>&1 || { echo "pop plain failed"; tail -3 gpurun_out/pop_plain_${tag}.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1000 --csv --log-file gpurun_out/pop_launches_${tag}.csv $PCMD > gpurun_out/pop_ncu1_${tag}.log 2>&1
echo "pop launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"k_merge_deadtime|k_background$|k_trig_bin|k_trig_eval|k_sample_energy_env|k_source_thin|k_gap_prefix|k_energy_tables|k_thin_nodes" \
    -c 27 -f -o gpurun_out/prof_pop_${tag} $PCMD > gpurun_out/pop_ncu2_${tag}.log 2>&1
echo "pop full rc=$?"
post gpurun_out/prof_pop_${tag}.ncu-rep pop_${tag} k_merge_deadtime 'k_background$' k_trig_bin k_trig_eval k_sample_energy_env k_source_thin
tail -2 gpurun_out/pop_plain_${tag}.log
fi
