"""CPU study for the next round: lane utilisation of the blend kernels on the bench scene.
Takes the north-star scene (131072 Gaussians, 512x512 camera), evaluates every (pixel, splat) pair of a
128x128 crop at the image centre (the busy region) with the reference's predicates (no early termination:
at this opacity distribution almost no pixel saturates) and reports, per (8x4 warp block, splat) visit,
how many lanes are valid -- what a warp-synchronous visit pays for vs what it uses.
    python tools/visit_stats.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_cull_invariant as t
from lara_b200 import scene as S
from oracle import oracle as O

f32 = np.float32
P, H, W = 131072, 512, 512
sc = S.scene(P, 0)
cam = S.cameras(8, H, W, 0)[0]
run = O.run_scene(sc, cam, torch.ones(3))
vis = np.asarray(run.radii) > 0
T = np.asarray(run.transMat).astype(f32); c = np.asarray(run.center).astype(f32); rad = np.asarray(run.radii)
o = sc["opacities"].numpy().reshape(-1).astype(f32)
x0 = y0 = 192; n = 128
touch = vis & (o >= 1 / 255.0) & (c[:, 0] + rad > x0) & (c[:, 0] - rad < x0 + n) & (c[:, 1] + rad > y0) & (c[:, 1] - rad < y0 + n)
idx = np.nonzero(touch)[0]
print("Gaussians touching the crop:", len(idx))
hist = np.zeros(33, np.int64)
pairs = 0
for s in range(0, len(idx), 2000):
    ii = idx[s:s + 2000]
    Ts = T[ii].copy(); cs = c[ii].copy()
    # shift pixel coordinates: evaluate only the crop by offsetting the homography (pix = crop + (x0, y0))
    # k = pix.x*Tw - Tu  with pix.x = x' + x0  ->  Tu' = Tu - x0*Tw ; same for Tv, y0
    Ts[:, 0:3] -= f32(x0) * Ts[:, 6:9]; Ts[:, 3:6] -= f32(y0) * Ts[:, 6:9]
    cs[:, 0] -= x0; cs[:, 1] -= y0
    valid, _, _ = t._valid_pairs(Ts, cs, o[ii], n, n)
    v = valid.reshape(len(ii), n // 4, 4, n // 8, 8).sum(axis=(2, 4))       # valid lanes per 8x4 block
    pairs += int(valid.sum())
    hist += np.bincount(v.reshape(-1), minlength=33)
visits = hist[1:].sum()
print(f"valid pairs {pairs}, warp visits with >=1 valid lane {visits}, mean valid lanes/visit {pairs / visits:.1f}")
cum = 0
for k in (1, 2, 4, 8, 12, 16, 24, 32):
    sel = hist[1:k + 1].sum(); lanes = (np.arange(1, k + 1) * hist[1:k + 1]).sum()
    print(f"  visits with <= {k:2d} valid lanes: {100 * sel / visits:5.1f} % of visits, {100 * lanes / pairs:5.1f} % of the pairs")
