#!/bin/bash
# usage: scripts/gpu_scale.sh <N> [tag]: the driver's N-GPU launch of bench.py (torchrun, one rank per GPU), full default legs
N=$1; tag=${2:-scale}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_${tag}_n$N.err | tail -3
python scripts/show_bench.py gpurun_out/bench_${tag}_n$N.json | grep -v "^   k_\|^      k_" | cut -c1-300
