"""The consumer: LaRa's lightning/renderer_2dgs.py must run unchanged on top of this package.

* CPU (only where /root/reference exists, i.e. in the build container): the reference's
  renderer_2dgs.py is imported *as is* with this repo's `diff_surfel_rasterization` on the path and
  builds a rasterizer from a MiniCam-like camera.
* GPU: the call pattern of Renderer.render_img (renderer_2dgs.py:167-268) -- activations,
  zeros+0 screenspace tensor with retain_grad, keyword call, the torch post-processing of the aux
  maps and a loss -- is replayed with this package and with the reference build; images and the
  gradients wrt the *raw* network outputs must agree.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from helpers import rel_err

REF_RENDERER = "/root/reference/lightning/renderer_2dgs.py"


@pytest.mark.skipif(not os.path.isfile(REF_RENDERER), reason="/root/reference not present (GPU box)")
def test_reference_renderer_imports_unchanged_on_cpu():
    import diff_surfel_rasterization as DSR
    assert "lara_b200" in DSR.GaussianRasterizer.__module__
    spec = importlib.util.spec_from_file_location("ref_renderer_2dgs", REF_RENDERER)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                      # does `from diff_surfel_rasterization import ...`
    assert mod.GaussianRasterizer is DSR.GaussianRasterizer
    r = mod.Renderer(sh_degree=1, white_background=True)

    class Cam:
        FoVx = FoVy = 0.75
        image_height = image_width = 32
        world_view_transform = torch.eye(4)
        full_proj_transform = torch.eye(4)
        camera_center = torch.zeros(3)
    rast = r.set_rasterizer(Cam(), device="cpu")
    assert isinstance(rast, DSR.GaussianRasterizer)
    assert rast.raster_settings.sh_degree == 1 and rast.raster_settings.image_width == 32


def _render_img_like(mod, cam, raw, bg, dev):
    """Replays renderer_2dgs.Renderer.render_img with module `mod` as the rasterizer package."""
    from lara_b200 import scene as S
    rs = S.settings_for(cam, bg, 1, dev, mod.GaussianRasterizationSettings)
    # MiniCam hands over w2c.transpose(0, 1): a NON-contiguous view matrix (lightning/utils.py:40)
    w2c = rs.viewmatrix.t().contiguous()
    rs = rs._replace(viewmatrix=w2c.transpose(0, 1))
    assert not rs.viewmatrix.is_contiguous()
    rast = mod.GaussianRasterizer(raster_settings=rs)
    opacity = torch.sigmoid(raw["opacity"])
    scales = torch.exp(raw["scales"])
    rotations = torch.nn.functional.normalize(raw["rotations"])
    centers = raw["centers"]
    screenspace = torch.zeros_like(centers, dtype=centers.dtype, requires_grad=True, device=dev) + 0
    screenspace.retain_grad()
    img, radii, allmap = rast(means3D=centers, means2D=screenspace, shs=raw["shs"], opacities=opacity,
                              scales=scales, rotations=rotations, cov3D_precomp=None)
    img = img.clamp(0, 1)
    alpha = allmap[1:2]
    normal = (allmap[2:5].permute(1, 2, 0) @ (rs.viewmatrix[:3, :3].T)).permute(2, 0, 1)
    # (LaRa divides by alpha and nan_to_num()s the result; with anomaly mode on, 0/0 at empty pixels
    #  would trip DivBackward itself, so the replay guards the denominator)
    depth = torch.nan_to_num(allmap[0:1] / alpha.clamp_min(1e-6), 0, 0)
    dist = allmap[6:7]
    loss = ((img - 0.3) ** 2).mean() + 0.2 * (normal ** 2).mean() + 1000.0 * dist.mean() + 0.1 * depth.mean() + alpha.mean()
    return img, allmap, radii, loss, screenspace


@pytest.mark.gpu
def test_render_img_call_pattern_matches_reference(cuda_device, reference):
    import diff_surfel_rasterization as DSR
    from lara_b200 import scene as S
    dev = cuda_device
    P = 20000
    sc = S.scene(P, 11)
    cam = S.cameras(2, 160, 160, 3)[1]
    bg = torch.ones(3)
    base = {"centers": sc["means3D"], "shs": sc["shs"], "opacity": torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)),
            "scales": torch.log(sc["scales"]), "rotations": sc["rotations"] * 1.7}
    res = []
    for mod in (DSR, reference):
        raw = {k: v.to(dev).clone().requires_grad_(True) for k, v in base.items()}
        with torch.autograd.set_detect_anomaly(True):
            img, allmap, radii, loss, ss = _render_img_like(mod, cam, raw, bg, dev)
            loss.backward()
        res.append((img.detach().cpu().numpy(), allmap.detach().cpu().numpy(), radii.cpu().numpy(),
                    {k: v.grad.cpu().numpy() for k, v in raw.items()}, ss.grad.cpu().numpy(), float(loss)))
    a, b = res
    assert np.array_equal(a[0].view(np.int32), b[0].view(np.int32))
    assert np.array_equal(a[1].view(np.int32), b[1].view(np.int32))
    assert np.array_equal(a[2], b[2])
    assert a[5] == b[5]
    for k in a[3]:
        assert np.isfinite(a[3][k]).all()
        assert rel_err(a[3][k], b[3][k]) < 1e-4, k
    assert rel_err(a[4], b[4]) < 1e-4        # viewspace (means2D) gradient used for densification statistics
