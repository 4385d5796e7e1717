"""Write profiles/README.md from the committed evidence files (bench lines, ncu summaries, launch shares)."""
import json, os

P = "profiles"


def jl(name):
    p = os.path.join(P, name)
    if not os.path.isfile(p):
        return None
    txt = open(p).read()
    lines = [l for l in txt.split("\n") if l.startswith("{")]
    return json.loads(lines[-1]) if lines else json.load(open(p))


b1v7 = jl("bench_r02_n1_v7.json")
b1, ref, b2, b4, b8 = jl("bench_r02_n1.json"), jl("bench_r02_reference_n1.json"), jl("bench_r02_n2.json"), jl("bench_r02_n4.json"), jl("bench_r02_n8.json")
ncu1 = json.load(open(os.path.join(P, "ncu_summary_r02.json")))
ncu8 = json.load(open(os.path.join(P, "ncu_summary_r02_views8.json")))
rv = ref["value"]
o = []
w = o.append
w("# profiles/ — measured evidence, round 2\n")
w("All numbers: B200 (gpurun boxes), SM clock 1965 MHz with no throttle reasons during the runs (`clocks` in the bench JSON), CUDA 12.9, "
  "workload = north-star point `scene(131072, seed 0)`, 512×512, SH degree 1, white background unless stated.  Round-1 files (`*_r01*`) are kept for comparison.\n")
w("| file | what |\n|---|---|")
w("| `bench_r02_n1.json`, `bench_r02_reference_n1.json`, `pytest_gpu_r02.log` | `python bench.py --steps 20 --warmup 5`, `--impl reference --steps 5 --warmup 3` and `pytest -m gpu` (69 passed); the candidate line and the test log are of the final kernels (`tools/r2_final.sh`, blend-backward variant 18 = the default) |")
w("| `bwd_variants_r02.log`, `parity_sweep_r02_final60.log` | A/B of the blend-backward variants in one process (`tools/time_kernels.py --variants`), and `tools/parity_sweep.py 7 60` on the final default: 60/60 bit-exact, worst gradient relative error 1.5e-5 |")
w("| `bench_r02_n2.json`, `bench_r02_n4.json`, `bench_r02_n8.json` | `bench.py --steps 10 --warmup 3` under `torch.distributed.run` on 2 / 4 / 8 B200s of one box (`tools/r2_gpu_n.sh`) — taken BEFORE the last blend-backward change (variant 7, 318 µs per view); `bench_r02_n1_v7.json` is the 1-GPU line of that same kernel, the scaling table below is computed against it |")
w("| `ncu_summary_r02.{json,md}` | one `ncu --set full --clock-control none --import-source on` capture per kernel, ONE view per launch (the drop-in path; `tools/profile_view.py`) |")
w("| `ncu_summary_r02_views8.{json,md}` | the same for the batched launch set, 8 views per launch (`tools/time_kernels.py --steps 1`) |")
w("| `launches_bench_r02.csv`, `launch_shares_r02.json` | `ncu --metrics gpu__time_duration.sum --clock-control none -c 400` over `bench.py --steps 2 --warmup 1` (cold-cache, serialised: compare shares) |")
w("| `roofline_traffic.json` | DRAM bytes and warp instructions per view of the dominant kernel (from the batched capture); `bench.py` reports them as `roofline.traffic` / `roofline.issue` |")
w("| `sass_histogram_r02.md` | per-kernel SASS opcode histogram of the built library (`tools/sass_histogram.py`): sm_100a cubins, `FFMA2/FMUL2/FADD2`, `REDG.E.ADD.F32x4`, `LDG.E.128`; no spills except 12 bytes in the default blend backward (96 registers for five CTAs per SM; stored and reloaded once per 128-splat round, outside the pair loops) |")
w("| `parity_sweep_r02.log` | `python tools/parity_sweep.py 7 250` on the kernels as of the scaling runs (blend-backward variant 7): 250 random configurations (500–300k Gaussians, 64–768 px incl. ragged sizes, SH 0–3, needles, saturated scenes, cameras inside the cloud) — **250/250 bit-exact** on every integer state array, colour and aux maps; worst gradient relative error 2.2e-5 |")
w("| `sanitizer_r02.txt` | compute-sanitizer racecheck / memcheck / synccheck (parity, batched-views, loss, decoder-layout tests): 0 hazards, 0 errors — run before the last blend-backward change (variant 7); variant 18 keeps the same two `__syncwarp()` points around the X tile and was not re-run under the sanitizer (GPU budget) |")
w("| `allreduce_probe_n8.log` | latency of the step's one collective (11.5 MB all-reduce) on 8 GPUs under a few NCCL settings |\n")

w("## Headline (1×B200)\n")
w("| | views/s (fwd+bwd) | vs reference |\n|---|---|---|")
w(f"| reference CUDA build, its own Python API, inputs resident (`--impl reference`) | {rv:.0f} | 1.00× |")
w(f"| candidate `value`: all 8 views of a scene in ONE launch set (`srf_views_*`), inputs resident | {b1['value']:.0f} | {b1['value']/rv:.2f}× |")
e = b1["e2e"]
w(f"| candidate `e2e`: the unchanged per-view drop-in API + autograd, pinned host params+cameras in, gradients out, every step | {e['value']:.0f} | {e['value']/rv:.2f}× |")
if "dropin_api_inputs_resident" in e:
    w(f"| the same per-view API with resident inputs (like the reference arm) | {e['dropin_api_inputs_resident']['value']:.0f} | {e['dropin_api_inputs_resident']['value']/rv:.2f}× |")
w(f"| `e2e.batched`: host in / host out through the batched public entry | {e['batched']['value']:.0f} | {e['batched']['value']/rv:.2f}× |")
w(f"| sustained ({b1['extra']['sustained']['seconds']:.1f} s back to back, no L2 flush) | {b1['extra']['sustained']['value']:.0f} | — |")
w(f"| CPU oracle port, {b1.get('cpu_baseline', {}).get('cores', '?')} host cores (reported baseline, not a target) | {b1.get('cpu_baseline', {}).get('value', 0):.2f} | — |\n")
w("Round 1 → round 2 on the same definitions: `value` 1727 → %.0f, `e2e` 1315 → %.0f (the per-view path still pays one launch set per view).\n" % (b1["value"], e["value"]))

w("## BASELINE configs and scaling (driver-visible `extra` blocks of the bench line)\n")
w("| config | candidate | reference (1 B200) | ratio |\n|---|---|---|---|")
for k, label in (("C2_32k_512_1view", "C2: 32 768 Gaussians, 512², 1 view (ms per fwd+bwd view)"),
                 ("C4_256k_1024_4views_per_gpu", "C4 share: 262 144 Gaussians, 1024², 4 views per GPU (views/s)"),
                 ("C3_raster_share_524k_8views_per_scene", "C3 rasterizer share: 8 scenes × 8 views at 524 288 Gaussians (ms per step)"),
                 ("strong_8_global_views", "north-star literal: 8 views in total (views/s)")):
    a, r = b1["extra"].get(k), ref["extra"].get(k)
    if not a or not r:
        continue
    if "ms per" in label:
        av, rvv = a["ms_per_step"], r["ms_per_step"]
        extra = f" (drop-in API: {a['dropin_api_ms_per_view']:.3f} ms)" if "dropin_api_ms_per_view" in a else ""
        w(f"| {label} | {av:.3f}{extra} | {rvv:.3f} | {rvv/av:.2f}× |")
    else:
        w(f"| {label} | {a['value']:.0f} | {r['value']:.0f} | {a['value']/r['value']:.2f}× |")
w("")
w("Scaling runs (blend-backward variant 7 on every row, so that the efficiency column compares like with like; the GPU budget of the round ended before they could be repeated with variant 18, which shortens every rank's step by the same 0.44 ms and leaves the ~0.15 ms of collective + rank skew at 8 GPUs unchanged — expected efficiency there 0.96):\n")
w("| GPUs (8 views per GPU, weak scaling) | `value` views/s | efficiency | `e2e` | `e2e.batched` | collective µs (incl. rank skew) | all-reduced gradients vs single-rank sum | 8 views in total (strong) |\n|---|---|---|---|---|---|---|---|")
for n, b in ((1, b1v7), (2, b2), (4, b4), (8, b8)):
    if not b:
        continue
    st = b["extra"].get("strong_8_global_views", {})
    gc = b.get("grad_check_detail") or {}
    w(f"| {n} | {b['value']:.0f} | {b['value']/(n*b1v7['value']):.3f} | {b['e2e']['value']:.0f} | {b['e2e']['batched']['value']:.0f} | "
      f"{((b.get('collective_us') or {}).get('rank0', 0) if isinstance(b.get('collective_us'), dict) else (b.get('collective_us') or 0)):.0f} | {gc.get('status', '—')} {('(%.1e)' % gc['max_rel_err']) if gc else ''} | {st.get('value', 0):.0f} ({st.get('ms_per_step', 0):.2f} ms/step) |")
w("")
w("The first 4-GPU run of the final tree measured 7494 views/s (0.930) with rank 2 about 7 % behind the others in `rank_compute_ms` (`bench_r02_n4_run1.json`); the rerun on a fresh box is the row above — the difference is the box, not the code.\n")
w("The pure all-reduce takes ~100–120 µs on 8 GPUs (`allreduce_probe_n8.log`, ranks in lock-step); the rest of `collective_us` in the "
  "bench is ranks waiting for the slowest one.  Reference at 8 GPUs: it is single-GPU, so 8×B200 vs 1×B200 is "
  f"{(b8['value']/rv if b8 else 0):.1f}× (`value`) / {(b8['e2e']['value']/rv if b8 else 0):.1f}× (`e2e`).\n")

w("## Per-kernel time per view (live CUDA events inside `bench.py`, 8 views per launch) and ncu\n")
k = b1["kernels_us"]
names = ["render_bwd", "render_fwd", "preprocess_fwd", "scatter", "preprocess_bwd", "sort_small", "tile_scan", "sort_big"]
r1 = {"render_bwd": 403, "render_fwd": 185, "preprocess_fwd": 20.7, "scatter": 22.6, "preprocess_bwd": 18.8, "sort_small": 13.9, "tile_scan": 10.8, "sort_big": 6.6}


def find(lst, name):
    for d in lst:
        if name in d["kernel"]:
            return d
    return None


w("| kernel | µs/view batched (round 1, per-view launches) | ncu, one view per launch: µs / warp instr / issue active / DRAM GB/s | ncu, 8 views per launch: µs per view / DRAM GB/s |\n|---|---|---|---|")
for n in names:
    a, c = find(ncu1, n + "_kernel"), find(ncu8, n + "_kernel")
    s1 = f"{a['time_us']:.1f} / {a['warp_instructions']/1e6:.1f} M / {a['issue_active_pct']:.0f} % / {a.get('dram_gbs', 0):.0f}" if a else "—"
    s8 = f"{c['time_us']/8:.1f} / {c.get('dram_gbs', 0):.0f}" if c else "—"
    w(f"| `{n}_kernel` | {k[n]:.1f} ({r1[n]}) | {s1} | {s8} |")
tot = sum(k.values())
w(f"| sum | {tot:.0f} (681) | | |\n")
ro = b1["roofline"]
rt = json.load(open(os.path.join(P, "roofline_traffic.json")))
w("## Roofline statement\n")
w(f"`roofline` in the bench line is computed as specified: SURVEY 8d algorithmic bytes for K7, `148·R_eff + 64·Npix` per view × the views of a launch = "
  f"{ro['algorithmic_bytes_per_launch']/1e6:.0f} MB, ÷ the live launch duration {ro['avg_launch_us']:.0f} µs, ÷ the measured {ro['peak']:.0f} GB/s: "
  f"**{ro['achieved']:.0f} GB/s, frac {ro['frac']:.3f}**; ncu's DRAM traffic for that launch is {(ro['traffic'] or 0)/1e6:.0f} MB (below the algorithmic bytes: the records and "
  "index lists stay in the 126 MB L2).  The kernel is bound by instruction issue" +
  (f" — `roofline.issue`: {rt['render_bwd_warp_instructions_per_view']*8/1e6:.0f} M warp instructions per launch (ncu, this kernel) = {rt['render_bwd_warp_instructions_per_view']*8/ro['issue']['issue_slots']:.2f} of the issue slots of 148 SMs × 4 schedulers "
   f"(the `roofline.issue` block inside `bench_r02_n1.json` was written with the previous kernel's count, {ro['issue']['warp_instructions_per_launch']/1e6:.0f} M, still in `roofline_traffic.json` at that moment; `bench.py` now ignores counts captured for another kernel variant)" if ro.get("issue") else "") +
  ".  Round 1 → round 2: 308 M → 231 M warp instructions per view in the backward (two-phase kernel; 186 M with the phase-2 lanes allotted in proportion to work), 128 M → 118 M in the forward (cull + compaction before the masks).  "
  "By the same §8(d) accounting the streaming kernels sit at: K1 0.28, binning (scan + scatter + sorts) 0.90, K8+K9 1.27 of the measured HBM roof "
  "(DESIGN.md §3 table); K1 is latency-bound on its per-tile counting atomics (ncu: long-scoreboard stalls 12 per issue).\n")
sv = os.path.join(P, "scene_views_r02.log")
if os.path.isfile(sv):
    w("## One scene-step of the renderer: 8 target views of one Gaussian set + loss + backward, 512×512 (`tools/bench_scene_views.py`)\n")
    w("The loop of `lightning/network.py:486-495`, the cat of `:525`, the `loss.py` terms (MSE + 1000·distortion + 0.2·normal consistency), `loss.backward()` (measured with blend-backward variant 7; the final default shortens every candidate row by a further ~0.44 ms per scene-step at 131 072 Gaussians):\n")
    w("```")
    for line in open(sv):
        if line.startswith("P="):
            w(line.rstrip()[:170])
    w("```\n")
open(os.path.join(P, "README.md"), "w").write("\n".join(o) + "\n")
print("\n".join(o)[:3000])
