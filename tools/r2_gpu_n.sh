#!/bin/bash
# multi-GPU runs of round 2: bash tools/r2_gpu_n.sh N [probe]
N=$1
mkdir -p gpurun_out
if [ "$2" == "probe" ]; then
  run() { env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tools/allreduce_probe.py 2>&1 | grep -E "^\{|rror" ; }
  { run NCCL_DEBUG=WARN; run NCCL_ALGO=Ring; run NCCL_ALGO=Tree; run NCCL_ALGO=NVLS; run NCCL_PROTO=LL128; run NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=32; run NCCL_NVLS_ENABLE=0; } > gpurun_out/allreduce_probe_n$N.log 2>&1
  cat gpurun_out/allreduce_probe_n$N.log
  exit 0
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err
echo "rc=$?"; tail -5 gpurun_out/bench_r02_n$N.err; cat gpurun_out/bench_r02_n$N.json | cut -c1-4000
