# On an 8-GPU box:  bash tools/scaling_run.sh   -> gpurun_out/scale_*.json
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n "$@" 2>/dev/null | tail -1; }
run 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n4.json
run 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n8.json
run 8 --steps 5 --warmup 3 --no-cpu-baseline --P 262144 --size 1024 --views 4 > gpurun_out/scale_c4_n8.json
for f in gpurun_out/scale_n4.json gpurun_out/scale_n8.json gpurun_out/scale_c4_n8.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['n_gpus'], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['config']['workload'][:60])"; done
