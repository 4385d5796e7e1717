"""The batched launch set (srf_views_*: every kernel carries a view dimension) against the per-view path:
outputs and internal state bit-identical, summed parameter gradients equal to the sum of per-view gradients;
plus the transMat_precomp + SH path of the per-Gaussian backward."""
import numpy as np
import pytest
import torch

from helpers import rel_err, to_dev

pytestmark = pytest.mark.gpu

KEYS = ("means3D", "shs", "opacities", "scales", "rotations")


@pytest.mark.parametrize("P,H,W,V,deg", [(20000, 256, 256, 3, 1), (60000, 300, 500, 5, 3), (131072, 512, 512, 8, 1)])
def test_batched_views_equal_per_view_calls(cuda_device, P, H, W, V, deg):
    from lara_b200 import rasterizer as R, scene as S
    from lara_b200.debug import unpack_state
    dev = cuda_device
    sc = S.scene(P, 7, sh_degree=deg)
    scd = to_dev(sc, dev)
    cams = S.cameras(V, H, W, 3)
    bgs = [torch.full((3,), b) for b in ([1.0, 0.0, 0.5] * V)[:V]]      # per-view backgrounds (network.py:489-490)
    sets = [S.settings_for(c, b, deg, dev, R.GaussianRasterizationSettings) for c, b in zip(cams, bgs)]
    ups = [tuple(t.to(dev) for t in S.upstream_grads(H, W, i)) for i in range(V)]

    packed = R.pack_cameras(sets, dev)
    color, allmap, radii, st = R.forward_views_raw(scd["means3D"], scd["shs"], None, scd["opacities"], scd["scales"],
                                                   scd["rotations"], None, packed, sets[0].tanfovx, sets[0].tanfovy, H, W, deg)
    gcs = torch.stack([u[0] for u in ups]).contiguous()
    gas = torch.stack([u[1] for u in ups]).contiguous()
    g = R.backward_views_raw(st, radii, scd["means3D"], scd["shs"], None, scd["scales"], scd["rotations"], None, packed,
                             sets[0].tanfovx, sets[0].tanfovy, H, W, deg, gcs, gas, need_means2D=True)
    torch.cuda.synchronize()
    counts = st.resolve()

    sums = None
    for v in range(V):
        c1, a1, r1, s1 = R.forward_raw(scd["means3D"], scd["shs"], None, scd["opacities"], scd["scales"], scd["rotations"],
                                       None, sets[v])
        g1 = R.backward_raw(s1, r1, scd["means3D"], scd["shs"], None, scd["scales"], scd["rotations"], None, sets[v],
                            ups[v][0], ups[v][1])
        torch.cuda.synchronize()
        assert torch.equal(color[v], c1) and torch.equal(allmap[v], a1) and torch.equal(radii[v], r1), v
        assert counts[v] == s1.num_rendered
        ub, u1 = unpack_state(st, P, H, W, view=v), unpack_state(s1, P, H, W)
        for k in ("ranges", "point_list", "n_contrib", "accum"):
            assert torch.equal(ub[k], u1[k]), (v, k)
        g1 = {k: t.double() for k, t in g1.items() if t is not None}
        sums = g1 if sums is None else {k: sums[k] + g1[k] for k in sums}
    for k in ("means3D", "sh", "opacities", "scales", "rotations", "means2D"):
        assert rel_err(g[k].cpu().numpy(), sums[k].cpu().numpy()) < 2e-5, k


def test_render_views_single_launch_set(cuda_device):
    """sharded.render_views issues one launch per kernel for all views of the call (<= 10 launches per scene-step)."""
    from lara_b200 import _lib, rasterizer as R, scene as S, sharded
    dev = cuda_device
    P, H, W, V = 30000, 256, 256, 6
    sc = S.scene(P, 1)
    params = {k: sc[k].to(dev) for k in KEYS}
    sets = [S.settings_for(c, torch.ones(3), 1, dev, R.GaussianRasterizationSettings) for c in S.cameras(V, H, W, 1)]
    up = tuple(t.to(dev) for t in S.upstream_grads(H, W, 0))
    _lib.profile_begin()
    sharded.render_views(params, sets, lambda vid, c, a: up)
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    assert all(n == 1 for _, n in prof.values()), prof
    assert sum(n for _, n in prof.values()) <= 10


@pytest.mark.parametrize("deg,use_sh", [(1, True), (0, True), (1, False)])
def test_transmat_precomp_backward_with_sh(cuda_device, deg, use_sh):
    """cov3D_precomp (= transMat_precomp) together with SHs: the SH vjp must run (reference backward.cu:596),
    dL_dsh / dL_dmeans3D equal the scale/rotation path's SH-only part; no NaN, no fault."""
    from lara_b200 import rasterizer as R, scene as S
    from lara_b200.debug import unpack_state
    dev = cuda_device
    P, H, W = 5000, 128, 128
    sc = S.scene(P, 4, sh_degree=deg)
    scd = to_dev(sc, dev)
    cam = S.cameras(1, H, W, 4)[0]
    st = S.settings_for(cam, torch.ones(3), deg, dev, R.GaussianRasterizationSettings)
    gc, ga = [t.to(dev) for t in S.upstream_grads(H, W, 4)]
    # reference run through scales / rotations to obtain the homographies it builds
    c0, a0, r0, s0 = R.forward_raw(scd["means3D"], scd["shs"], None, scd["opacities"], scd["scales"], scd["rotations"], None, st)
    torch.cuda.synchronize()
    T = unpack_state(s0, P, H, W)["transMat"].contiguous().clone()
    T[r0 <= 0] = 0
    g0 = R.backward_raw(s0, r0, scd["means3D"], scd["shs"], None, scd["scales"], scd["rotations"], None, st, gc, ga)
    colors = None
    if not use_sh:
        colors = unpack_state(s0, P, H, W)["rgb"].contiguous().clone()
    c1, a1, r1, s1 = R.forward_raw(scd["means3D"], scd["shs"] if use_sh else None, colors, scd["opacities"], None, None, T, st)
    g1 = R.backward_raw(s1, r1, scd["means3D"], scd["shs"] if use_sh else None, colors, None, None, T, st, gc, ga)
    torch.cuda.synchronize()
    assert torch.equal(r1, r0) and torch.equal(c1, c0)
    for k, t in g1.items():
        if t is not None:
            assert torch.isfinite(t).all(), k
    assert g1["cov3Ds_precomp"] is not None and float(g1["cov3Ds_precomp"].abs().max()) > 0
    if use_sh:
        assert rel_err(g1["sh"].cpu().numpy(), g0["sh"].cpu().numpy()) < 1e-5
        if deg > 0:
            assert float(g1["means3D"].abs().max()) > 0          # the view-direction term of the SH vjp
    else:
        assert g1["sh"] is None and rel_err(g1["colors_precomp"].cpu().numpy(),
                                            R.backward_raw(s0, r0, scd["means3D"], None, colors, scd["scales"], scd["rotations"],
                                                           None, st, gc, ga)["colors_precomp"].cpu().numpy()) < 1e-5
