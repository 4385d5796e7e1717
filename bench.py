#!/usr/bin/env python
"""bench.py -- fwd+bwd views/sec of the surfel-rasterizer hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json north_star point): synthetic ``scene(131072, seed 0)``,
512x512, white background, degree-1 SH, LaRa-like upstream gradients; every rank renders
`--views` (default 8) target views per step, forward AND backward.  View-sharded weak
scaling: N ranks -> N*views distinct views per step over the same Gaussian set, parameter
gradients accumulated in one flat buffer per rank and summed with a single NCCL
all-reduce per step.  A "step" = those fwd+bwd views + the all-reduce.

Printed JSON (one line, rank 0):
  value   : views/s, whole job, inputs resident in HBM, through lara_b200.sharded.render_views
            (raw C-ABI calls; per-step CUDA events on the launching stream, L2 flushed between
            steps outside the timed spans, max over ranks)
  e2e     : the same metric through the reference-facing drop-in API
            (diff_surfel_rasterization.GaussianRasterizer + autograd), with the Gaussian
            parameters and cameras copied from pinned host memory every step and the summed
            parameter gradients read back to the host every step
  roofline: dominant kernel (render_bwd) -- SURVEY 8d algorithmic bytes per launch / its live
            CUDA-event duration (srf_profile_*), against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline: the CPU oracle port (oracle/surfel_oracle.c, OpenMP) on one view of the same
            workload (N=1, rank 0 only)
--impl reference times the reference's own CUDA build (oracle/_ref) through its own Python
API on one GPU (rank 0), same workload; the reference has no CPU rasterizer.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "rasterizer fwd+bwd views/sec at {P} Gaussians x {S}x{S}"      # formatted with --P / --size (defaults: the north-star point)
UNIT = "views/s"
KERNELS_PER_VIEW = 8  # preprocess_fwd, tile_scan, scatter, sort_small, sort_big, render_fwd, render_bwd, preprocess_bwd


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--views", type=int, default=8, help="target views per GPU per step")
    ap.add_argument("--P", type=int, default=131072)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=3, help="CUDA streams the views of a step are spread over")
    ap.add_argument("--skip-value", action="store_true", help="diagnostic: skip the resident-input timing loop")
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (rank 0)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def build_workload(P, size, views_total, dev):
    from lara_b200 import scene as S
    sc = S.scene(P, 0, sh_degree=1)
    cams = S.cameras(views_total, size, size, 0)
    gc, ga = S.upstream_grads(size, size, 0, lara_like=True)
    return sc, cams, gc, ga


def timed_steps(step_fn, steps, warmup, flush, world, dev):
    """W untimed + K timed steps; per-step CUDA events, L2 flush between steps (untimed)."""
    import torch.distributed as dist
    for _ in range(warmup):
        flush.zero_()
        step_fn()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    t0 = time.perf_counter()
    for i in range(steps):
        flush.zero_()               # evict the previous step's working set from the 126 MB L2
        starts[i].record()
        step_fn()
        ends[i].record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, wall


def cpu_baseline(P, size):
    """Oracle port on the host cores, one fwd+bwd view of the same workload."""
    from lara_b200 import scene as S
    from oracle import oracle as O
    sc = S.scene(P, 0, sh_degree=1)
    cam = S.cameras(8, size, size, 0)[0]
    gc, ga = S.upstream_grads(size, size, 0, lara_like=True)
    O.load()
    t0 = time.perf_counter()
    run = O.run_scene(sc, cam, torch.ones(3))
    run.backward(gc, ga)
    dt = time.perf_counter() - t0
    run.close()
    return {"value": 1.0 / dt, "unit": UNIT, "cores": O.threads(), "kind": "port",
            "sample": f"1 view fwd+bwd of the {P}-Gaussian {size}x{size} workload, CPU oracle (C + OpenMP), {dt:.1f} s"}


def run_reference(args, rank, local, world):
    """Reference arm: the unmodified reference CUDA build through its own Python API, rank 0 only."""
    if rank != 0:
        return
    from oracle import ref as REF
    from lara_b200 import scene as S
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    base = {"impl": "reference", "metric": METRIC.format(P=args.P, S=args.size), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic"}
    if not REF.available():
        # the oracle always exists: fall back to the CPU port
        cb = cpu_baseline(args.P, args.size)
        cb["kind"] = "port"
        base.update({"value": cb["value"], "ms_per_step": 1e3 / cb["value"], "cpu_baseline": cb,
                     "config": {"workload": f"scene({args.P},seed0) {args.size}x{args.size}, 1 view/step on host cores"},
                     "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base), flush=True)
        return
    ref = REF.load()
    sc, cams, gc, ga = build_workload(args.P, args.size, args.views, dev)
    scd = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in sc.items()}
    gc, ga = gc.to(dev), ga.to(dev)
    bg = torch.ones(3)
    sets = [S.settings_for(c, bg, 1, dev, ref.GaussianRasterizationSettings) for c in cams]
    leaves = {k: scd[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)

    def step():
        for v in leaves.values():
            v.grad = None
        for rs in sets:
            rast = ref.GaussianRasterizer(raster_settings=rs)
            c, rd, am = rast(means3D=leaves["means3D"], means2D=m2d, shs=leaves["shs"], opacities=leaves["opacities"],
                             scales=leaves["scales"], rotations=leaves["rotations"])
            torch.autograd.backward((c, am), (gc, ga))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    sampler.start()
    ms, wall = timed_steps(step, args.steps, 0, flush, 1, dev)
    clocks = sampler.stop()
    value = args.views * args.steps / (ms / 1e3)
    base.update({
        "value": value, "ms_per_step": ms / args.steps,
        "config": {"workload": f"scene({args.P},seed0) {args.size}x{args.size} sh1 white bg, {args.views} views/step fwd+bwd, reference CUDA build on 1 B200",
                   "views_per_gpu": args.views, "l2": "flushed between steps (256 MiB write, untimed)"},
        "clocks": clocks,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                         "sample": "reference's own CUDA rasterizer (oracle/_ref) on 1 B200 -- the reference ships no CPU rasterizer"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })
    print(json.dumps(base), flush=True)


def main():
    args = parse_args()
    rank, local, world = dist_info()
    if args.impl == "reference":
        run_reference(args, rank, local, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the rasterizer has no CPU path)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=dev)

    from lara_b200 import scene as S
    from lara_b200 import rasterizer as R
    from lara_b200 import sharded, _lib
    from lara_b200.debug import unpack_state
    import diff_surfel_rasterization as DSR   # the drop-in name LaRa imports

    _lib.load()
    V = args.views
    total_views = V * world
    sc, cams, gc_h, ga_h = build_workload(args.P, args.size, total_views, dev)
    my_ids = sharded.shard_views(total_views, rank, world)
    bg = torch.ones(3)
    params = {k: sc[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    gc, ga = gc_h.to(dev), ga_h.to(dev)
    my_sets = [S.settings_for(cams[i], bg, 1, dev, R.GaussianRasterizationSettings) for i in my_ids]
    P, M = args.P, int(params["shs"].shape[1])
    grads = sharded.GradBuffer(P, M, dev)

    def upstream(vid, color, allmap):
        return gc, ga

    def step():
        grads.zero_()
        sharded.render_views(params, my_sets, upstream, grads=grads, view_ids=my_ids, streams=args.streams)
        grads.all_reduce()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(0 if args.skip_value else args.warmup):
        flush.zero_(); step()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.start()
    if args.skip_value:
        ms, wall = 1.0, 0.0
    else:
        ms, wall = timed_steps(step, args.steps, 0, flush, world, dev)
    value = total_views * args.steps / (ms / 1e3)

    # ---- end to end through the reference-facing API, host buffers in / gradients out
    pinned = {k: sc[k].pin_memory() for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    # all cameras of the step travel in one pinned buffer: [V, 16 view | 16 proj | 3 campos | pad]
    cam_host = torch.zeros((len(my_ids), 36), dtype=torch.float32)
    for r_, i in enumerate(my_ids):
        cam_host[r_, 0:16] = cams[i].viewmatrix.reshape(-1)
        cam_host[r_, 16:32] = cams[i].projmatrix.reshape(-1)
        cam_host[r_, 32:35] = cams[i].campos
    cam_host = cam_host.pin_memory()
    bg_dev = bg.to(dev)
    host_out = torch.empty(grads.flat.numel(), dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * 4 for t in pinned.values()) + cam_host.numel() * 4
    d2h = host_out.numel() * 4

    def step_e2e():
        dp = {k: v.to(dev, non_blocking=True).requires_grad_(True) for k, v in pinned.items()}
        m2d = torch.zeros_like(dp["means3D"], requires_grad=True)
        cam_dev = cam_host.to(dev, non_blocking=True)
        for r_, i in enumerate(my_ids):
            c = cams[i]
            rs = DSR.GaussianRasterizationSettings(
                image_height=c.image_height, image_width=c.image_width, tanfovx=c.tanfovx, tanfovy=c.tanfovy,
                bg=bg_dev, scale_modifier=1.0, viewmatrix=cam_dev[r_, 0:16].view(4, 4),
                projmatrix=cam_dev[r_, 16:32].view(4, 4), sh_degree=1, campos=cam_dev[r_, 32:35],
                prefiltered=False, debug=False)
            rast = DSR.GaussianRasterizer(raster_settings=rs)
            color, radii, allmap = rast(means3D=dp["means3D"], means2D=m2d, shs=dp["shs"], opacities=dp["opacities"],
                                        scales=dp["scales"], rotations=dp["rotations"])
            torch.autograd.backward((color, allmap), (gc, ga))
        flat = torch.cat([dp[k].grad.reshape(-1) for k in ("means3D", "shs", "opacities", "scales", "rotations")])
        if world > 1:
            dist.all_reduce(flat)
        host_out[:flat.numel()].copy_(flat, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()   # the caller reads the result

    e2e_steps = max(3, args.steps // 2)
    ms_e2e, _ = timed_steps(step_e2e, e2e_steps, 3, flush, world, dev)
    e2e_value = total_views * e2e_steps / (ms_e2e / 1e3)
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the dominant kernel (render_bwd), live CUDA-event timing of every launch
    roof = None
    kern = None
    if rank == 0:
        _lib.profile_begin()
    prof_steps = min(args.steps, 5)

    def step_single_stream():
        # per-kernel durations are only meaningful without concurrent kernels from other streams
        grads.zero_()
        sharded.render_views(params, my_sets, upstream, grads=grads, view_ids=my_ids, streams=1)
        grads.all_reduce()

    for _ in range(prof_steps):
        flush.zero_(); step_single_stream()
    torch.cuda.synchronize(dev)
    if rank == 0:
        kern = _lib.profile_end()
        # measured sizes of view 0 for the algorithmic byte count
        color, allmap, radii, st = R.forward_raw(params["means3D"], params["shs"], None, params["opacities"],
                                                 params["scales"], params["rotations"], None, my_sets[0])
        torch.cuda.synchronize(dev)
        u = unpack_state(st, P, args.size, args.size)
        gx = (args.size + 15) // 16
        ncon = u["n_contrib"][0].view(gx, 16, gx, 16).permute(0, 2, 1, 3).reshape(gx * gx, 256)
        r_eff = int(ncon.max(dim=1).values.sum().item())
        npix = args.size * args.size
        alg_bytes = 148 * r_eff + 64 * npix          # K7: (76+72)*R_eff + 64*Npix  
        tot_ms, n_launch = kern["render_bwd"]
        dur_s = tot_ms / 1e3 / max(n_launch, 1)
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                if args.P == 131072 and args.size == 512:      # the capture is of the default workload only
                    traffic = json.load(f).get("render_bwd_dram_bytes_per_launch")
        except Exception:
            pass
        achieved = alg_bytes / dur_s / 1e9
        roof = {"kernel": "render_bwd_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "R": st.num_rendered, "R_eff": r_eff,
                "avg_launch_us": dur_s * 1e6,
                "note": "the blend kernels are fp32-issue bound, not HBM bound (DESIGN.md); per-kernel us in 'kernels_us'"}

    if rank == 0:
        out = {
            "metric": METRIC.format(P=args.P, S=args.size), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"scene({args.P},seed0) {args.size}x{args.size} sh1 white bg, {V} views/GPU/step fwd+bwd, "
                                   f"view-sharded over {world} GPU(s) + 1 NCCL all-reduce of param grads",
                       "views_per_gpu": V, "global_views_per_step": total_views,
                       "parallelism": f"view-shard x{world}", "streams_per_gpu": args.streams,
                       "l2": "flushed between steps (256 MiB write, untimed)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "diff_surfel_rasterization.GaussianRasterizer + autograd, pinned host in / grads out"},
            "gpu_launches": KERNELS_PER_VIEW * V * args.steps * world,
            "clocks": clocks,
            "roofline": roof,
            "kernels_us": {k: (v[0] * 1e3 / v[1] if v[1] else 0.0) for k, v in kern.items()} if kern else None,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.P, args.size)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
