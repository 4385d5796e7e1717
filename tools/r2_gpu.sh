#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_view.py --iters 40 --warmup 10 --P 131072
timeout 300 python tools/profile_view.py --iters 40 --warmup 10 --P 32768
timeout 300 python tools/profile_view.py --iters 40 --warmup 10 --P 524288
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'resident dropin',round(d['e2e']['dropin_api_inputs_resident']['value'],1),'batched e2e',round(d['e2e']['batched']['value'],1))"
