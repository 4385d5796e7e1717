# round-1 final evidence refresh (run on the GPU box):  bash tools/refresh_evidence.sh
mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/bench_ref_r01b.err | tail -1 > gpurun_out/bench_ref_r01b.json
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r01b.err | tail -1 > gpurun_out/bench_r01b.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r01b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/gpu_check.py > gpurun_out/check_r01b.log 2>&1
python tools/parity_sweep.py 7 250 > gpurun_out/parity_sweep_r01b.log 2>&1
python tools/bench_render_img.py 131072 > gpurun_out/render_img_r01b.log 2>&1
python tools/bench_render_img.py 524288 >> gpurun_out/render_img_r01b.log 2>&1
python tools/bench_scene_views.py 131072 > gpurun_out/scene_views_r01b.log 2>&1
python tools/bench_scene_views.py 524288 >> gpurun_out/scene_views_r01b.log 2>&1
(for tool in memcheck racecheck synccheck initcheck; do echo "== compute-sanitizer --tool $tool"; timeout 900 compute-sanitizer --tool $tool python tools/profile_view.py --P 20000 --size 160 --iters 1 --warmup 0 2>&1 | grep -E "ERROR SUMMARY|mine P=|Error|error" | head -8; done) > gpurun_out/sanitizer_r01b.txt 2>&1
tail -2 gpurun_out/parity_sweep_r01b.log; cat gpurun_out/sanitizer_r01b.txt; tail -3 gpurun_out/check_r01b.log
