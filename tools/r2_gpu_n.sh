#!/bin/bash
# multi-GPU runs of round 2: bash tools/r2_gpu_n.sh N [probe|skew]
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$2" == "probe" ]; then
  run() { env "$@" timeout 300 $TR --master-port 29577 tools/allreduce_probe.py 2>&1 | grep -E "^\{|rror" ; }
  { run NCCL_DEBUG=WARN; run NCCL_ALGO=Ring; run NCCL_PROTO=LL128; run NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=32; run NCCL_NVLS_ENABLE=0; } > gpurun_out/allreduce_probe_n$N.log 2>&1
  cat gpurun_out/allreduce_probe_n$N.log
  exit 0
fi
if [ "$2" == "skew" ]; then
  for spin in 1 0; do      # sleeping vs spinning wait for the instance-count event
    SRF_BLOCKING_EVENT_WAIT=$spin timeout 600 $TR --master-port 2953$spin bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('blocking_wait=$spin value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'coll',d['collective_us'],'compute',d['rank_compute_ms'],'spread',d['rank_step_ms'])"
  done
  nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv,noheader
  exit 0
fi
timeout 900 $TR --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err
echo "rc=$?"; tail -5 gpurun_out/bench_r02_n$N.err; cat gpurun_out/bench_r02_n$N.json | cut -c1-4000
