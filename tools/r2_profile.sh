#!/bin/bash
# final evidence of round 2: one ncu --set full capture per kernel (one view, drop-in path), the launch list of the
# bench command, compute-sanitizer on the small configurations.  Outputs under gpurun_out/ (summarised into profiles/).
mkdir -p gpurun_out
for k in render_bwd_kernel render_fwd_kernel preprocess_fwd_kernel preprocess_bwd_kernel scatter_kernel sort_small_kernel tile_scan_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_$k python tools/profile_view.py --iters 1 --warmup 2 > gpurun_out/ncu_$k.log 2>&1
done
# batched launch set (8 views per launch): the two blend kernels and the two per-Gaussian kernels
for k in render_bwd_kernel render_fwd_kernel preprocess_fwd_kernel preprocess_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none -k regex:$k -s 3 -c 1 -f -o gpurun_out/r02_views8_$k python tools/time_kernels.py --steps 1 > gpurun_out/ncu_v8_$k.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/bench_under_ncu_r02.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_views.py tests/test_loss.py tests/test_decoder_layout.py -m gpu -q -x -k "not 131072" > gpurun_out/sanitizer_r02.txt 2>&1
tail -5 gpurun_out/sanitizer_r02.txt
ls -la gpurun_out/*.ncu-rep | tail -12
