#!/bin/bash
# A/B of the blend-backward variants in one process, then parity + bench + ncu with the fastest one.
#   bash tools/r2_exp.sh "7,9,12,16,17" [validated-variant]
mkdir -p gpurun_out
V=${1:-7,9,10,11,12,13,14,15,7,9,11}
L=gpurun_out/bwd_variants_r02c.log
timeout 300 python tools/time_kernels.py --variants $V > $L 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/bwd_variants_r02c.log"):
    if l.startswith("{"):
        d = json.loads(l); print(d["bwd_variant"], d["us_per_view"]["render_bwd"], d["views_per_s"])
    else:
        print(l.strip()[:200])
PY
best=$(grep '^BEST' $L | awk '{print $2}')
[ -z "$best" ] && { echo "no timing"; exit 1; }
echo "best variant $best"
[ "$best" == "$2" ] && { echo "already validated"; exit 0; }
export SRF_BWD_VARIANT=$best
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v$best.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_v$best.log
timeout 200 python tools/parity_sweep.py 7 60 > gpurun_out/parity_sweep_v$best.log 2>&1; echo "sweep rc=$?"; tail -1 gpurun_out/parity_sweep_v$best.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_v$best.json 2> gpurun_out/bench_r02_v$best.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r02_v$best.json
