"""Which splat of a pixel's tile list sits at the alpha = 1/255 threshold, is it outside its cull octagon, and how
edge-on is it?    python tools/diag_pixel.py P size seed x y [x y ...]"""
import math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lara_b200 import scene as S, rasterizer as R
from lara_b200.debug import unpack_state
from helpers import to_dev

P, size, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pts = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(4, len(sys.argv), 2)]
dev = torch.device("cuda:0")
sc = S.scene(P, seed, sh_degree=1)
cam = S.cameras(3, size, size, seed)[seed % 3]
scd = to_dev(sc, dev)
st = S.settings_for(cam, torch.ones(3), 1, dev, R.GaussianRasterizationSettings)
color, allmap, radii, state = R.forward_raw(scd["means3D"], scd["shs"], None, scd["opacities"], scd["scales"], scd["rotations"], None, st)
torch.cuda.synchronize()
u = {k: v.cpu().numpy() for k, v in unpack_state(state, P, size, size).items()}
rec = u["rec"].astype(np.float64)
oct16 = u["rec"][:, 20:24].copy().view(np.float16).astype(np.float64)      # lo/hi pairs along x, y, x+y, x-y
gx = (size + 15) // 16
fx = size / (2 * cam.tanfovx)
for (x, y) in pts:
    tile = (y // 16) * gx + x // 16
    s, e = u["ranges"][tile]
    ids = u["point_list"][s:e]
    px, py = x + 0.5, y + 0.5
    print(f"pixel ({x},{y}) tile {tile} list {e - s} n_contrib {u['n_contrib'][0][y, x]}")
    for pos, i in enumerate(ids):
        r = rec[i]
        Tu = np.array([r[0], r[2], r[4]]); Tv = np.array([r[1], r[3], r[5]]); Tw = np.array([r[6], r[7], r[8]])
        cx, cy, opac = r[9], r[10], r[11]
        k = px * Tw - Tu; l = py * Tw - Tv
        p = np.cross(k, l)
        if p[2] == 0:
            continue
        sx, sy = p[0] / p[2], p[1] / p[2]
        rho3d = sx * sx + sy * sy
        rho2d = 2 * ((cx - px) ** 2 + (cy - py) ** 2)
        alpha = min(0.99, opac * math.exp(-0.5 * min(rho3d, rho2d)))
        if abs(alpha * 255 - 1) < 0.03:
            o = oct16[i]
            inside = (cx + o[0] <= px <= cx + o[1]) and (cy + o[2] <= py <= cy + o[3]) and \
                     (cx + cy + o[4] <= px + py <= cx + cy + o[5]) and (cx - cy + o[6] <= px - py <= cx - cy + o[7])
            n = r[12:15]
            pv = np.array([(cx - size / 2) / fx, (cy - size / 2) / fx, 1.0]); pv /= np.linalg.norm(pv)
            kappa = (abs(k[0] * l[1]) + abs(k[1] * l[0])) / abs(p[2])
            print(f"   pos {pos} id {i} alpha*255 {alpha * 255:.5f} rho3d {rho3d:.4f} rho2d {rho2d:.4f} inside_octagon {inside} "
                  f"cos {float(np.dot(n, pv)):+.5f} kappa {kappa:.1f} opac {opac:.4f} radius {radii[i].item()} "
                  f"oct x[{o[0]:.2f},{o[1]:.2f}] y[{o[2]:.2f},{o[3]:.2f}] d=({px - cx:.2f},{py - cy:.2f})")
