#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/parity_sweep.py 7 250 > gpurun_out/parity_sweep_r02.log 2>&1; tail -2 gpurun_out/parity_sweep_r02.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cpu_oracle" > gpurun_out/racecheck_r02.txt 2>&1; tail -4 gpurun_out/racecheck_r02.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_views.py tests/test_loss.py tests/test_decoder_layout.py -m gpu -q -x -k "cpu_oracle or 20000 or loss or decoder" > gpurun_out/sanitizer_r02.txt 2>&1; tail -4 gpurun_out/sanitizer_r02.txt
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cpu_oracle" > gpurun_out/synccheck_r02.txt 2>&1; tail -3 gpurun_out/synccheck_r02.txt
