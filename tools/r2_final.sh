#!/bin/bash
# final evidence of round 2 on the final tree: A/B of the last two blend-backward variants, then the -m gpu suite,
# a 60-configuration parity sweep, the bench line, one ncu --set full capture of the blend backward and the launch
# list of the bench command.
mkdir -p gpurun_out
L=gpurun_out/bwd_variants_r02d.log
timeout 120 python tools/time_kernels.py --variants 17,18,17,18 > $L 2>&1
python - <<'PY'
import json
t = {}
for l in open("gpurun_out/bwd_variants_r02d.log"):
    if l.startswith("{"):
        d = json.loads(l); t.setdefault(d["bwd_variant"], []).append(d["us_per_view"]["render_bwd"]); print(d["bwd_variant"], d["us_per_view"]["render_bwd"], d["views_per_s"])
a, b = min(t.get(17, [1e9])), min(t.get(18, [1e9]))
open("gpurun_out/final_variant.txt", "w").write("18" if b < 0.996 * a else "17")
PY
best=$(cat gpurun_out/final_variant.txt)
echo "variant for the evidence: $best"
[ "$best" != "17" ] && export SRF_BWD_VARIANT=$best
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r02.log
timeout 120 python tools/parity_sweep.py 7 60 > gpurun_out/parity_sweep_r02_final60.log 2>&1; echo "sweep rc=$?"; tail -1 gpurun_out/parity_sweep_r02_final60.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_r02_n1.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:render_bwd_kernel -s 2 -c 1 -f -o gpurun_out/r02_render_bwd_kernel python tools/profile_view.py --iters 1 --warmup 2 > gpurun_out/ncu_render_bwd_kernel.log 2>&1; echo "ncu rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/bench_under_ncu_r02.log 2>&1; echo "launch list rc=$?"
timeout 100 ncu --set full --clock-control none -k regex:render_bwd_kernel -s 3 -c 1 -f -o gpurun_out/r02_views8_render_bwd_kernel python tools/time_kernels.py --steps 1 > gpurun_out/ncu_v8_render_bwd_kernel.log 2>&1; echo "ncu views8 rc=$?"
