"""Diagnose a forward/backward mismatch against the reference build on one configuration.
    python tools/diag_parity.py P size seed [cam_index]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lara_b200 import scene as S
from oracle import ref as REF
from helpers import run_candidate, to_dev, rel_err

P, size, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
ref = REF.load()
sc = S.scene(P, seed, sh_degree=1)
cam = S.cameras(3, size, size, seed)[seed % 3]
bg = torch.ones(3)
gc, ga = S.upstream_grads(size, size, seed)
mine = run_candidate(sc, cam, bg, dev, grads=(gc, ga))
scd = to_dev(sc, dev)
st = S.settings_for(cam, bg, 1, dev, ref.GaussianRasterizationSettings)
r = REF.forward_raw(ref, scd, st)
r = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in r.items()}
print("R", mine["num_rendered"], r["num_rendered"], "n_contrib equal", np.array_equal(mine["n_contrib"][0], r["n_contrib"][0]))
for k in ("color", "allmap", "accum"):
    a, b = mine[k], r[k]
    ne = a.view(np.int32) != b.view(np.int32)
    print(k, "differing values", int(ne.sum()), "of", ne.size, "max abs diff", float(np.abs(a - b).max()))
    if ne.any():
        idx = np.argwhere(ne)
        d = np.abs(a - b)[ne]
        order = np.argsort(-d)[:8]
        for o in order:
            c, y, x = idx[o]
            print("   ch", c, "x", x, "y", y, "mine", a[c, y, x], "ref", b[c, y, x], "n_contrib", mine["n_contrib"][0][y, x], r["n_contrib"][0][y, x],
                  "T", mine["accum"][0][y, x], r["accum"][0][y, x])
        ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))[ne]
        print("   ulp distance: median", float(np.median(ulp)), "max", int(ulp.max()), "pixels", len(set(map(tuple, idx[:, 1:]))))
