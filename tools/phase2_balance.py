"""CPU study: lane balance of the two phases of the blend backward on the bench scene.
For every 8x4 warp block of a 128x128 crop of the north-star view, takes the splats with at least one valid pixel
in the block (in list order), forms the groups of 16 the kernel forms, and compares the phase-2 trip count of
   two lanes per splat sharing its pixels (max_i ceil(c_i / 2)),
   lanes allotted per splat in proportion to its pixels (smallest C with sum_i ceil(c_i / C) <= 32),
   the bound ceil(sum_i c_i / 32),
and the phase-1 trip count (max over the block's pixels of the pixel's contributing splats in the group) for groups of 16
and of 32 splats.
    python tools/phase2_balance.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_cull_invariant as t
from lara_b200 import scene as S
from oracle import oracle as O

f32 = np.float32
P, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 512, 512
sc = S.scene(P, 0)
cam = S.cameras(8, H, W, 0)[0]
run = O.run_scene(sc, cam, torch.ones(3))
vis = np.asarray(run.radii) > 0
T = np.asarray(run.transMat).astype(f32); c = np.asarray(run.center).astype(f32); rad = np.asarray(run.radii)
o = sc["opacities"].numpy().reshape(-1).astype(f32)
x0 = y0 = 192; n = 128
touch = vis & (o >= 1 / 255.0) & (c[:, 0] + rad > x0) & (c[:, 0] - rad < x0 + n) & (c[:, 1] + rad > y0) & (c[:, 1] - rad < y0 + n)
idx = np.nonzero(touch)[0]
nb = (n // 4) * (n // 8)
per_block = [[] for _ in range(nb)]
per_block_px = [[] for _ in range(nb)]      # per block: list of [32] bool rows, one per splat with a valid pixel in it
for s in range(0, len(idx), 2000):
    ii = idx[s:s + 2000]
    Ts = T[ii].copy(); cs = c[ii].copy()
    Ts[:, 0:3] -= f32(x0) * Ts[:, 6:9]; Ts[:, 3:6] -= f32(y0) * Ts[:, 6:9]
    cs[:, 0] -= x0; cs[:, 1] -= y0
    valid, _, _ = t._valid_pairs(Ts, cs, o[ii], n, n)
    v = valid.reshape(len(ii), n // 4, 4, n // 8, 8).sum(axis=(2, 4)).reshape(len(ii), nb)
    g, b = np.nonzero(v)
    vp = valid.reshape(len(ii), n // 4, 4, n // 8, 8).transpose(0, 1, 3, 2, 4).reshape(len(ii), nb, 32)
    for gi, bi in zip(g, b):
        per_block[bi].append(int(v[gi, bi]))
        per_block_px[bi].append(vp[gi, bi])
tot = {"pairs": 0, "groups": 0, "two": 0, "prop": 0, "bound": 0, "lanes_prop": 0}
for lst in per_block:
    for k in range(0, len(lst), 16):
        cnt = np.array(lst[k:k + 16])
        Tsum = int(cnt.sum())
        two = int(np.ceil(cnt / 2).max())
        C = next(C for C in range(1, 17) if int(np.ceil(cnt / C).sum()) <= 32)
        tot["pairs"] += Tsum; tot["groups"] += 1; tot["two"] += two; tot["prop"] += C; tot["bound"] += -(-Tsum // 32)
        tot["lanes_prop"] += int(np.ceil(cnt / C).sum())
g = tot["groups"]
print(f"P={P}: groups {g}, pairs/group {tot['pairs'] / g:.1f}")
print(f"  trips/group  two lanes per splat {tot['two'] / g:.2f}   proportional {tot['prop'] / g:.2f}   bound {tot['bound'] / g:.2f}")
print(f"  useful lanes per trip: two {tot['pairs'] / tot['two']:.1f}  proportional {tot['pairs'] / tot['prop']:.1f}; lanes allotted {tot['lanes_prop'] / g:.1f}")
for G in (16, 32):
    trips = pairs = groups = 0
    for rows in per_block_px:
        for k in range(0, len(rows), G):
            m = np.stack(rows[k:k + G])                  # [splats in group, 32 pixels]
            trips += int(m.sum(axis=0).max()); pairs += int(m.sum()); groups += 1
    print(f"  phase 1, groups of {G}: {trips / groups:.2f} trips/group, {pairs / trips:.1f} of 32 lanes contributing per trip "
          f"({trips / 1e3:.0f} k trips in the crop)")
