#!/bin/bash
# multi-GPU bench of round 2: bash tools/r2_gpu_n.sh N
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err
echo "rc=$?"; tail -5 gpurun_out/bench_r02_n$N.err; cat gpurun_out/bench_r02_n$N.json | cut -c1-4000
