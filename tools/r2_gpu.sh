#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_pixel.py 262144 1024 2 685 578 698 857 > gpurun_out/diag_pixel.log 2>&1
tail -12 gpurun_out/diag_pixel.log
rm -f gpurun_out/time_kernels.log
for v in 2 3 4 5; do
  SRF_BWD_VARIANT=$v timeout 300 python tools/time_kernels.py >> gpurun_out/time_kernels.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/time_kernels.log'):
    try:
        d = json.loads(l); print(d['bwd_variant'], d['us_per_view']['render_bwd'], d['us_per_view']['render_fwd'], d['views_per_s'])
    except Exception: print(l[:200])
PY
SRF_BWD_VARIANT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd_kernel -s 2 -c 1 -f -o gpurun_out/r2b_render_bwd python tools/profile_view.py --iters 1 --warmup 2 > gpurun_out/ncu_bwd.log 2>&1
SRF_BWD_VARIANT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_fwd_kernel -s 2 -c 1 -f -o gpurun_out/r2b_preprocess_fwd python tools/profile_view.py --iters 1 --warmup 2 > gpurun_out/ncu_k1.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
