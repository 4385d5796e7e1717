"""GPU-side development check: candidate vs the reference build (oracle/_ref) stage by stage,
plus a quick fwd+bwd timing of both.  Writes gpurun_out/check.json.

    gpurun -- python tools/gpu_check.py [--quick]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lara_b200 import scene as S  # noqa: E402
from lara_b200 import rasterizer as R  # noqa: E402
from lara_b200.debug import unpack_state  # noqa: E402
from oracle import ref as REF  # noqa: E402


def rel(a, b):
    a = a.double(); b = b.double()
    d = (a - b).abs().max().item() if a.numel() else 0.0
    m = b.abs().max().item() if b.numel() else 0.0
    return d / m if m > 0 else d


def bits_equal(a, b):
    return int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum().item())


def to_dev(sc, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in sc.items()}


def run_case(P, H, W, seed, deg, bgv, ref, dev, timing=True):
    sc = to_dev(S.scene(P, seed, sh_degree=deg), dev)
    cam = S.cameras(3, H, W, seed)[seed % 3]
    bg = torch.full((3,), bgv, dtype=torch.float32)
    res = {"P": P, "H": H, "W": W, "seed": seed, "deg": deg, "bg": bgv}

    mine_set = S.settings_for(cam, bg, deg, dev, R.GaussianRasterizationSettings)
    color, allmap, radii, st = R.forward_raw(sc["means3D"], sc["shs"], None, sc["opacities"], sc["scales"],
                                             sc["rotations"], None, mine_set)
    torch.cuda.synchronize()
    mine = unpack_state(st, P, H, W)
    res["num_rendered"] = st.num_rendered
    if ref is not None:
        ref_set = S.settings_for(cam, bg, deg, dev, ref.GaussianRasterizationSettings)
        r = REF.forward_raw(ref, sc, ref_set)
        torch.cuda.synchronize()
        vis = r["radii"] > 0
        res["ref_num_rendered"] = r["num_rendered"]
        res["radii_mismatch"] = int((radii != r["radii"]).sum().item())
        res["tiles_touched_mismatch"] = int((mine["tiles_touched"] != r["tiles_touched"]).sum().item())
        res["visible"] = int(vis.sum().item())
        res["depth_bits_mismatch"] = bits_equal(mine["depths"][vis], r["depths"][vis])
        res["transMat_bits_mismatch"] = bits_equal(mine["transMat"][vis], r["transMat"][vis])
        res["means2D_bits_mismatch"] = bits_equal(mine["means2D"][vis], r["means2D"][vis])
        res["normal_bits_mismatch"] = bits_equal(mine["normal"][vis], r["normal_opacity"][vis][:, :3])
        res["rgb_rel"] = rel(mine["rgb"][vis], r["rgb"][vis])
        res["rgb_bits_mismatch"] = bits_equal(mine["rgb"][vis], r["rgb"][vis])
        if st.num_rendered == r["num_rendered"]:
            res["point_list_mismatch"] = int((mine["point_list"] != r["point_list"]).sum().item())
        res["ranges_mismatch"] = int((mine["ranges"] != r["ranges"]).sum().item())
        res["n_contrib_mismatch"] = int((mine["n_contrib"][0] != r["n_contrib"][0]).sum().item())
        # the reference leaves the median plane uninitialised for tiles with an empty range
        gx = (W + 15) // 16
        nonempty = (r["ranges"][:, 1] > r["ranges"][:, 0])
        ty = torch.arange(H, device=dev) // 16
        tx = torch.arange(W, device=dev) // 16
        pixmask = nonempty[(ty[:, None] * gx + tx[None, :])]
        res["median_contrib_mismatch"] = int(((mine["n_contrib"][1] != r["n_contrib"][1]) & pixmask).sum().item())
        res["accum_bits_mismatch"] = bits_equal(mine["accum"], r["accum"])
        res["color_bits_mismatch"] = bits_equal(color, r["color"])
        res["allmap_bits_mismatch"] = bits_equal(allmap, r["allmap"])
        res["color_rel"] = rel(color, r["color"])
        res["allmap_rel"] = [rel(allmap[c], r["allmap"][c]) for c in range(8)]

    # backward through the public autograd surface of both
    gc, ga = S.upstream_grads(H, W, seed)
    gc, ga = gc.to(dev), ga.to(dev)

    def fwd_bwd(mod, settings):
        leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        rast = mod.GaussianRasterizer(raster_settings=settings)
        c, rd, am = rast(means3D=leaves["means3D"], means2D=m2d, shs=leaves["shs"], opacities=leaves["opacities"],
                         scales=leaves["scales"], rotations=leaves["rotations"])
        torch.autograd.backward((c, am), (gc, ga))
        g = {k: v.grad for k, v in leaves.items()}
        g["means2D"] = m2d.grad
        return g

    gm = fwd_bwd(R, mine_set)
    torch.cuda.synchronize()
    res["grad_finite"] = all(bool(torch.isfinite(v).all().item()) for v in gm.values())
    if ref is not None:
        gr = fwd_bwd(ref, ref_set)
        gr2 = fwd_bwd(ref, ref_set)
        torch.cuda.synchronize()
        res["grad_rel"] = {k: rel(gm[k], gr[k]) for k in gm}
        res["grad_noise_floor"] = {k: rel(gr2[k], gr[k]) for k in gm}

    if timing:
        def bench(mod, settings, iters=30, warm=10):
            leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
            m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)

            def step():
                rast = mod.GaussianRasterizer(raster_settings=settings)
                c, rd, am = rast(means3D=leaves["means3D"], means2D=m2d, shs=leaves["shs"], opacities=leaves["opacities"],
                                 scales=leaves["scales"], rotations=leaves["rotations"])
                torch.autograd.backward((c, am), (gc, ga))
            for _ in range(warm):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                step()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        res["ms_mine"] = bench(R, mine_set)
        if ref is not None:
            res["ms_ref"] = bench(ref, ref_set)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ref = REF.load() if REF.available() else None
    print("reference available:", ref is not None, flush=True)
    cases = [(4096, 128, 128, 0, 1, 1.0), (20000, 300, 500, 1, 3, 0.5), (32768, 512, 512, 0, 1, 1.0)]
    if not args.quick:
        cases += [(131072, 512, 512, 0, 1, 1.0), (524288, 512, 512, 2, 1, 0.0), (262144, 1024, 1024, 0, 1, 1.0)]
    out = []
    for c in cases:
        t = time.time()
        try:
            r = run_case(*c, ref, dev)
        except Exception as ex:  # keep going: one broken case must not hide the others
            import traceback
            traceback.print_exc()
            r = {"case": c, "error": repr(ex)}
        r["wall_s"] = time.time() - t
        print(json.dumps(r), flush=True)
        out.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "check.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
