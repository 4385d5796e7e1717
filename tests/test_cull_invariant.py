"""CPU check of the invariant the blend kernels' cull rests on:

    a (pixel, splat) pair outside the splat's conservative screen octagon can never reach
    alpha >= 1/255, so skipping it changes no result.

The octagon is restated here in numpy exactly as lara_b200/csrc/preprocess.cu computes it (the
tau-ellipse's extent along x, y, x+y, x-y from the reference's AABB quadratic form with
diag(tau,tau,-1), the rho2d disc, the 1/16 px + 0.1 % margin, outward fp16 rounding) and the pair
evaluation as reference forward.cu:353-398 does it; the homographies / centres come from the CPU
oracle.  Brute force over every (pixel, splat) pair of small scenes, needle-like splats included.
The CUDA implementation of the same test is covered by the bit-exact GPU parity tests."""
import numpy as np
import pytest
import torch

from lara_b200 import scene as S
from oracle import oracle as O

f32 = np.float32


def _half_outward(lo, hi):
    """fp16 rounding towards -inf for lo, +inf for hi (cvt.rd / cvt.ru)."""
    with np.errstate(over="ignore", invalid="ignore"):
        l16 = lo.astype(np.float16); h16 = hi.astype(np.float16)
        l16 = np.where(l16.astype(f32) > lo, np.nextafter(l16, np.float16(-np.inf)), l16)
        h16 = np.where(h16.astype(f32) < hi, np.nextafter(h16, np.float16(np.inf)), h16)
    return l16.astype(f32), h16.astype(f32)


def _octagon(T, centre, opac):
    """[P,4] lo / hi offsets from the centre along x, y, x+y, x-y (preprocess.cu, the cull block)."""
    Tu, Tv, Tw = T[:, 0:3], T[:, 3:6], T[:, 6:9]
    P = T.shape[0]
    lo = np.full((P, 4), 3.0e38, f32); hi = np.full((P, 4), -3.0e38, f32)
    vis = opac >= f32(0.00392156862745098)
    with np.errstate(all="ignore"):
        tau = f32(2.0) * np.log(opac * f32(255.0)) * f32(1.0001) + f32(1.0e-3)
        r2 = np.sqrt(f32(0.5) * tau)
        qw = tau * (Tw[:, 0] ** 2 + Tw[:, 1] ** 2) - Tw[:, 2] ** 2
        iq = f32(1.0) / qw
        rows = [Tu, Tv, Tu + Tv, Tu - Tv]
        dc = [centre[:, 0], centre[:, 1], centre[:, 0] + centre[:, 1], centre[:, 0] - centre[:, 1]]
        dr = [r2, r2, f32(1.41421357) * r2, f32(1.41421357) * r2]
        mg = [f32(0.0625), f32(0.0625), f32(0.0884), f32(0.0884)]
        for k in range(4):
            # rows translated to the splat's screen centre (row - centre * Tw): in absolute pixel coordinates
            # c*c - (...) cancels catastrophically at 1024^2 (round-2 finding, see preprocess.cu)
            r = (rows[k] - dc[k][:, None] * Tw).astype(f32)
            c = (tau * (r[:, 0] * Tw[:, 0] + r[:, 1] * Tw[:, 1]) - r[:, 2] * Tw[:, 2]) * iq
            h = np.sqrt(np.maximum(f32(0.0), c * c - (tau * (r[:, 0] ** 2 + r[:, 1] ** 2) - r[:, 2] ** 2) * iq))
            l0 = np.minimum(-dr[k], c - h); h0 = np.maximum(dr[k], c + h)
            m = mg[k] + f32(1.0e-3) * (h0 - l0) + f32(4.0e-7) * (np.abs(rows[k][:, 2]) + np.abs(dc[k] * Tw[:, 2])) * np.abs(Tw[:, 2] * iq)
            bounded = vis & (qw < 0)
            lo[:, k] = np.where(bounded, l0 - m, np.where(vis, f32(-3.0e38), f32(3.0e38)))
            hi[:, k] = np.where(bounded, h0 + m, np.where(vis, f32(3.0e38), f32(-3.0e38)))
    return _half_outward(lo, hi)


def _valid_pairs(T, centre, opac, H, W, x0=0, y0=0):
    """[P,H,W] bool: the pair is blended by the reference (forward.cu:353-398 predicates); pixels of the
    H x W window whose first pixel is (x0, y0)."""
    ys, xs = np.meshgrid(np.arange(y0, y0 + H, dtype=f32) + f32(0.5), np.arange(x0, x0 + W, dtype=f32) + f32(0.5), indexing="ij")
    px, py = xs[None], ys[None]
    g = lambda i: T[:, i][:, None, None]
    with np.errstate(all="ignore"):
        kx, ky, kz = px * g(6) - g(0), px * g(7) - g(1), px * g(8) - g(2)
        lx, ly, lz = py * g(6) - g(3), py * g(7) - g(4), py * g(8) - g(5)
        pz = kx * ly - ky * lx
        sx = (ky * lz - kz * ly) / pz
        sy = (kz * lx - kx * lz) / pz
        rho3d = sx * sx + sy * sy
        dx, dy = centre[:, 0][:, None, None] - px, centre[:, 1][:, None, None] - py
        rho2d = f32(2.0) * (dx * dx + dy * dy)
        rho = np.minimum(rho3d, rho2d)
        depth = np.where(rho3d <= rho2d, sx * g(6) + sy * g(7) + g(8), g(8))
        alpha = np.minimum(f32(0.99), opac[:, None, None] * np.exp(f32(-0.5) * rho))
        # NaN rho (p.z == 0) makes every comparison false -> not valid, as in the kernels
        return (pz != 0) & ~(depth < f32(0.2)) & (rho >= 0) & (alpha >= f32(0.00392156862745098)), px, py


@pytest.mark.parametrize("seed,needles,scale", [(0, False, 1.0), (1, True, 1.0), (2, False, 3.0), (3, True, 0.4)])
def test_no_blended_pair_lies_outside_its_octagon(seed, needles, scale):
    H = W = 64
    P = 1500
    sc = S.scene(P, seed)
    sc["scales"] = sc["scales"] * 6.0 * scale                     # P is small: LaRa-like footprints in pixels
    if needles:
        sc["scales"][:, 1] *= 0.04                                # edge-on / needle-like conics
    cam = S.cameras(3, H, W, seed)[seed % 3]
    O.load()
    run = O.run_scene(sc, cam, torch.ones(3))
    vis = np.asarray(run.radii) > 0
    assert vis.sum() > P // 2
    T = np.asarray(run.transMat)[vis].astype(f32)
    centre = np.asarray(run.center)[vis].astype(f32)
    opac = sc["opacities"].numpy().reshape(-1)[vis].astype(f32)

    lo, hi = _octagon(T, centre, opac)
    valid, px, py = _valid_pairs(T, centre, opac, H, W)
    eps = f32(0.0009765625)                                       # SRF_MASK_EPS of octagon_pixel_mask()
    c = lambda a: a[:, None, None]
    coords = [px - c(centre[:, 0]), py - c(centre[:, 1]),
              (px + py) - c(centre[:, 0] + centre[:, 1]), (px - py) - c(centre[:, 0] - centre[:, 1])]
    inside = np.ones_like(valid)
    for k in range(4):
        inside &= (coords[k] >= c(lo[:, k]) - eps) & (coords[k] <= c(hi[:, k]) + eps)
    missed = valid & ~inside
    assert int(missed.sum()) == 0, f"{int(missed.sum())} blended pairs outside the cull octagon"
    # negative control: an octagon 1 px too small does lose blended pairs (the check has teeth)
    shrunk = np.ones_like(valid)
    for k in range(4):
        shrunk &= (coords[k] >= c(lo[:, k]) + f32(1.0)) & (coords[k] <= c(hi[:, k]) - f32(1.0))
    assert int((valid & ~shrunk).sum()) > 0
    # ... and the octagon is not vacuous: a fair share of what it keeps is really blended
    assert valid.sum() > 1000
    # (thin oblique ellipses are where an octagon is loosest: ~0.25 for the needle scenes, ~0.6 otherwise)
    assert valid.sum() / max(int(inside.sum()), 1) > (0.15 if needles else 0.35)


@pytest.mark.parametrize("x0,y0,seed", [(720, 730, 0), (700, 250, 1), (250, 740, 2)])
def test_octagon_is_conservative_far_from_the_origin_and_for_faint_splats(x0, y0, seed):
    """Windows of a 1024 x 1024 view (pixel coordinates ~1000, along the diagonals ~2000) with opacities just above
    the 1/255 threshold: tiny tau-ellipses, where an extent computed in absolute coordinates loses all its digits
    (round 2: two wrongly culled pairs in a 262 144-splat 1024^2 view)."""
    size, win = 1024, 64
    P = 60000
    sc = S.scene(P, seed)
    g = torch.Generator().manual_seed(seed)
    faint = torch.rand((P, 1), generator=g) * 0.03 + 0.0035                            # tau in [~0, 4]
    barely = (1.0 / 255.0) * (1.0 + torch.rand((P, 1), generator=g) * 0.08)            # tau in [0, 0.16]: extents of ~1 px
    sc["opacities"] = torch.where(torch.rand((P, 1), generator=g) < 0.5, faint, barely).float()
    sc["scales"] = sc["scales"] * 0.6
    cam = S.cameras(3, size, size, seed)[seed % 3]
    O.load()
    run = O.run_scene(sc, cam, torch.ones(3))
    centre_all = np.asarray(run.center).astype(f32)
    vis = (np.asarray(run.radii) > 0) & (centre_all[:, 0] > x0 - 10) & (centre_all[:, 0] < x0 + win + 10) & \
          (centre_all[:, 1] > y0 - 10) & (centre_all[:, 1] < y0 + win + 10)
    assert vis.sum() > 100
    T = np.asarray(run.transMat)[vis].astype(f32)
    centre = centre_all[vis]
    opac = sc["opacities"].numpy().reshape(-1)[vis].astype(f32)
    lo, hi = _octagon(T, centre, opac)
    valid, px, py = _valid_pairs(T, centre, opac, win, win, x0, y0)
    eps = f32(0.0009765625)
    c = lambda a: a[:, None, None]
    coords = [px - c(centre[:, 0]), py - c(centre[:, 1]),
              (px + py) - c(centre[:, 0] + centre[:, 1]), (px - py) - c(centre[:, 0] - centre[:, 1])]
    inside = np.ones_like(valid)
    for k in range(4):
        inside &= (coords[k] >= c(lo[:, k]) - eps) & (coords[k] <= c(hi[:, k]) + eps)
    assert valid.sum() > 300
    assert int((valid & ~inside).sum()) == 0, f"{int((valid & ~inside).sum())} blended pairs outside the cull octagon"
