#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/time_kernels.log
for v in 2 8 7; do
  SRF_BWD_VARIANT=$v timeout 300 python tools/time_kernels.py >> gpurun_out/time_kernels.log 2>&1
done
SRF_BWD_VARIANT=8 timeout 300 python tools/time_kernels.py --P 524288 >> gpurun_out/time_kernels.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/time_kernels.log'):
    try:
        d = json.loads(l); print(d['P'], d['bwd_variant'], d['us_per_view']['render_bwd'], d['us_per_view']['render_fwd'], d['views_per_s'])
    except Exception: print(l[:200])
PY
for v in 8 7; do SRF_BWD_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_views.py tests/test_gpu_headline.py -m gpu -q -x 2>&1 | tail -2; done
