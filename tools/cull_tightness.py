"""CPU study for the next round: how tight are candidate cull shapes per pixel?
valid = pairs the reference blends; octagon = what the kernels use today (tests/test_cull_invariant.py);
obb = oriented box along the principal axes of the tau-ellipse (+ the rho2d disc), intersected with the octagon.
    python tools/cull_tightness.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_cull_invariant as t
from lara_b200 import scene as S
from oracle import oracle as O

f32 = np.float32


def obb_inside(T, centre, opac, px, py):
    """Extent of the tau-ellipse along its own principal directions (same quadratic form as the octagon)."""
    Tu, Tv, Tw = T[:, 0:3].astype(np.float64), T[:, 3:6].astype(np.float64), T[:, 6:9].astype(np.float64)
    tau = 2.0 * np.log(opac.astype(np.float64) * 255.0) * 1.0001 + 1e-3
    f = np.stack([tau, tau, -np.ones_like(tau)], 1)
    qw = (f * Tw * Tw).sum(1)
    def ext(r):      # centre and half-width of the ellipse's projection on the direction with row r
        c = (f * r * Tw).sum(1) / qw
        h = np.sqrt(np.maximum(0.0, c * c - (f * r * r).sum(1) / qw))
        return c, h
    cx, hx = ext(Tu); cy, hy = ext(Tv); cu, hu = ext(Tu + Tv)
    # covariance-like matrix of the ellipse from three directional half-widths: h(n)^2 = n^T S n
    sxx, syy = hx * hx, hy * hy
    sxy = (hu * hu - sxx - syy) / 2.0
    ang = 0.5 * np.arctan2(2 * sxy, sxx - syy)
    c1, s1 = np.cos(ang), np.sin(ang)
    r2 = np.sqrt(0.5 * tau)
    inside = np.ones(px.shape[:0] + (T.shape[0],) + px.shape[1:], bool)
    for (c, s) in ((c1, s1), (-s1, c1)):
        cc, hh = ext(c[:, None] * Tu + s[:, None] * Tv)
        d0 = c * centre[:, 0] + s * centre[:, 1]
        lo = np.minimum(cc - hh, d0 - r2) - 0.0625; hi = np.maximum(cc + hh, d0 + r2) + 0.0625
        proj = c[:, None, None] * px + s[:, None, None] * py
        inside &= (proj >= lo[:, None, None]) & (proj <= hi[:, None, None])
    return inside & (qw < 0)[:, None, None] | (qw >= 0)[:, None, None]


for name, needles, scale, P in (("bench-like", False, 1.0, 1500), ("large splats", False, 3.0, 1500), ("needles", True, 1.0, 1500), ("small needles", True, 0.4, 1500)):
    H = W = 64
    sc = S.scene(P, 5); sc["scales"] = sc["scales"] * 6.0 * scale
    if needles:
        sc["scales"][:, 1] *= 0.04
    cam = S.cameras(3, H, W, 5)[1]
    run = O.run_scene(sc, cam, torch.ones(3))
    vis = np.asarray(run.radii) > 0
    T = np.asarray(run.transMat)[vis].astype(f32); c = np.asarray(run.center)[vis].astype(f32)
    o = sc["opacities"].numpy().reshape(-1)[vis].astype(f32)
    keep = o >= 1 / 255.0
    T, c, o = T[keep], c[keep], o[keep]
    lo, hi = t._octagon(T, c, o)
    valid, px, py = t._valid_pairs(T, c, o, H, W)
    cc = lambda a: a[:, None, None]
    coords = [px - cc(c[:, 0]), py - cc(c[:, 1]), (px + py) - cc(c[:, 0] + c[:, 1]), (px - py) - cc(c[:, 0] - c[:, 1])]
    octa = np.ones_like(valid)
    for k in range(4):
        octa &= (coords[k] >= cc(lo[:, k])) & (coords[k] <= cc(hi[:, k]))
    aabb = np.ones_like(valid)
    for k in range(2):
        aabb &= (coords[k] >= cc(lo[:, k])) & (coords[k] <= cc(hi[:, k]))
    obb = obb_inside(T, c, o, px, py)
    both = octa & obb
    print(f"{name:14s} valid {int(valid.sum()):8d} | kept/valid: aabb {aabb.sum()/valid.sum():5.2f}  octagon {octa.sum()/valid.sum():5.2f}  "
          f"octagon+obb {both.sum()/valid.sum():5.2f} | missed by octagon+obb: {int((valid & ~both).sum())}")
