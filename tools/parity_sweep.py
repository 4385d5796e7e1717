"""Randomised bit-exact parity sweep against the reference build (oracle/_ref) on the GPU.
Varies P, image size (incl. ragged), SH degree, background, splat scale, opacity distribution,
camera distance/fov; checks every integer state array and colour / aux maps bit for bit and the
gradients to 1e-4 (reports the worst ratio to the reference's own noise)."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lara_b200 import scene as S
import diff_surfel_rasterization as DSR
from oracle import ref as REF
from helpers import run_candidate, to_dev, tile_pixel_mask, rel_err

dev = torch.device("cuda:0")
ref = REF.load()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
bad = 0
worst_grad = 0.0
t0 = time.time()
for it in range(N):
    P = int(rng.choice([500, 3000, 20000, 60000, 150000, 300000]))
    H = int(rng.choice([64, 100, 200, 256, 333, 512, 768])); W = int(rng.choice([64, 120, 200, 256, 400, 512, 700]))
    deg = int(rng.integers(0, 4)); bgv = float(rng.choice([0.0, 0.5, 1.0])); seed = int(rng.integers(0, 10000))
    sc = S.scene(P, seed, sh_degree=deg)
    sc["scales"] = sc["scales"] * float(rng.choice([0.3, 1.0, 2.5, 6.0]))
    if rng.random() < 0.3:
        sc["scales"][:, 1] *= 0.05                      # needle-like splats (edge-on conics)
    if rng.random() < 0.3:
        sc["opacities"] = torch.rand_like(sc["opacities"])          # many opaque splats -> saturation
    fov = float(rng.choice([0.4, 0.75, 1.3])); radius = float(rng.choice([0.9, 1.905, 4.0]))   # 0.9: camera inside the cloud
    cam = S.cameras(3, H, W, seed, fov=fov, radius=radius)[seed % 3]
    bg = torch.full((3,), bgv)
    gc, ga = S.upstream_grads(H, W, seed)
    mine = run_candidate(sc, cam, bg, dev, grads=(gc, ga))
    scd = to_dev(sc, dev)
    st = S.settings_for(cam, bg, deg, dev, ref.GaussianRasterizationSettings)
    r = REF.forward_raw(ref, scd, st)
    r = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in r.items()}
    vis = r["radii"] > 0
    errs = []
    def chk(name, ok):
        if not ok: errs.append(name)
    chk("radii", np.array_equal(mine["radii"], r["radii"]))
    chk("R", mine["num_rendered"] == r["num_rendered"])
    if mine["num_rendered"] == r["num_rendered"]:
        chk("point_list", np.array_equal(mine["point_list"], r["point_list"]))
    chk("ranges", np.array_equal(mine["ranges"], r["ranges"]))
    chk("n_contrib", np.array_equal(mine["n_contrib"][0], r["n_contrib"][0]))
    m = tile_pixel_mask(r["ranges"], H, W)
    chk("median", np.array_equal(mine["n_contrib"][1][m], r["n_contrib"][1][m]))
    chk("color", np.array_equal(mine["color"].view(np.int32), r["color"].view(np.int32)))
    chk("allmap", np.array_equal(mine["allmap"].view(np.int32), r["allmap"].view(np.int32)))
    chk("accum", np.array_equal(mine["accum"].view(np.int32), r["accum"].view(np.int32)))
    # gradients through the reference's autograd surface
    leaves = {k: scd[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    rast = ref.GaussianRasterizer(raster_settings=st)
    c_, rd_, am_ = rast(means3D=leaves["means3D"], means2D=torch.zeros_like(leaves["means3D"]), shs=leaves["shs"],
                        opacities=leaves["opacities"], scales=leaves["scales"], rotations=leaves["rotations"])
    torch.autograd.backward((c_, am_), (gc.to(dev), ga.to(dev)))
    for a_, b_ in (("g_means3D", "means3D"), ("g_sh", "shs"), ("g_opacities", "opacities"), ("g_scales", "scales"), ("g_rotations", "rotations")):
        gr = leaves[b_].grad.cpu().numpy()
        e = rel_err(mine[a_], gr)
        worst_grad = max(worst_grad, e)
        chk("grad_" + b_ + f"({e:.1e})", e < 1e-4 and np.isfinite(mine[a_]).all())
    tag = dict(P=P, H=H, W=W, deg=deg, bg=bgv, seed=seed, fov=fov, radius=radius, R=mine["num_rendered"], visible=int(vis.sum()))
    print(("FAIL " if errs else "ok   ") + json.dumps(tag) + (" " + ",".join(errs) if errs else ""), flush=True)
    bad += bool(errs)
print(f"{N - bad}/{N} configurations bit-exact; worst gradient rel err {worst_grad:.2e}; {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
