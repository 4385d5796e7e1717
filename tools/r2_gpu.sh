#!/bin/bash
# one gpurun call of round 2: parity tests + ncu captures of the blend kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd_kernel -s 2 -c 1 -f -o gpurun_out/r2a_render_bwd python tools/profile_view.py --iters 1 --warmup 2 > gpurun_out/ncu_bwd.log 2>&1
tail -3 gpurun_out/ncu_bwd.log
