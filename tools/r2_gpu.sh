#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_scene_views.py 524288 > gpurun_out/scene_views_r02.log 2>&1
timeout 600 python tools/bench_scene_views.py 131072 >> gpurun_out/scene_views_r02.log 2>&1
cat gpurun_out/scene_views_r02.log | cut -c1-200
