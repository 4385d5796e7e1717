#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
SRF_BWD_VARIANT=2 timeout 300 python tools/time_kernels.py 2>&1 | tail -1 | cut -c1-600
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02_n1.err; cat gpurun_out/bench_r02_n1.json | cut -c1-3500
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_r02_ref_n1.json 2> gpurun_out/bench_r02_ref_n1.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_r02_ref_n1.err; cat gpurun_out/bench_r02_ref_n1.json | cut -c1-1500
