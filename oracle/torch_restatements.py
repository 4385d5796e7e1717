"""Plain-torch restatements of reference code around the rasterizer, used ONLY as checkers by tests/ and
tools/ (oracle/ is test infrastructure; nothing under lara_b200/ imports it).  Each function cites the
reference lines it restates and is itself pinned against the reference run on the CPU (tests/test_epilogue.py,
tests/test_loss.py)."""
from __future__ import annotations

import torch


def render_img_epilogue_torch(rendered_image, allmap, rays, world_view_transform, depth_ratio=0.0, prex=""):
    """Plain-torch restatement of renderer_2dgs.py:220-268 (+ depth_to_normal :74-89).

    Test/measurement reference for the fused kernel -- this is what LaRa executes today."""
    rendered_image = rendered_image.clamp(0, 1)
    render_alpha = allmap[1:2]
    render_normal = allmap[2:5]
    render_normal = (render_normal.permute(1, 2, 0) @ (world_view_transform[:3, :3].T)).permute(2, 0, 1)
    render_depth_median = torch.nan_to_num(allmap[5:6], 0, 0)
    render_depth_expected = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
    render_dist = allmap[6:7]
    surf_depth = render_depth_expected * (1 - depth_ratio) + depth_ratio * render_depth_median
    points = (rays[..., :3].reshape(-1, 3) + surf_depth.reshape(-1, 1) * rays[..., 3:].reshape(-1, 3)).reshape(
        *surf_depth.shape[1:], 3)
    output = torch.zeros_like(points)
    dx = points[2:, 1:-1] - points[:-2, 1:-1]
    dy = points[1:-1, 2:] - points[1:-1, :-2]
    output[1:-1, 1:-1, :] = torch.nn.functional.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    surf_normal = output.permute(2, 0, 1) * render_alpha.detach()
    return {
        f"image{prex}": rendered_image.permute(1, 2, 0), f"depth{prex}": surf_depth.permute(1, 2, 0),
        f"acc_map{prex}": render_alpha.squeeze(0), f"rend_normal{prex}": render_normal.permute(1, 2, 0),
        f"depth_normal{prex}": surf_normal.permute(1, 2, 0), f"rend_dist{prex}": render_dist.squeeze(0),
    }


def lara_loss_torch(output, tar_rgb, iter, ssim=None):
    """Plain-torch restatement of lightning/loss.py:17-62 (``Losses.forward``) for the coarse outputs (prex '').

    output[k]: [B, H, V*W, C] (or [B, H, V*W]) as network.py:525/:531 builds them; tar_rgb: [B,V,H,W,3].
    ``ssim(img[B,3,H,VW], tar[B,3,H,VW])`` replaces pytorch_msssim's MS_SSIM (absent here); None drops the term."""
    B, V, H, W = tar_rgb.shape[:-1]
    tar = tar_rgb.permute(0, 2, 1, 3, 4).reshape(B, H, V * W, 3)
    stats = {}
    color_loss_all = (output["image"] - tar) ** 2
    loss = color_loss_all.mean()
    stats["mse"] = color_loss_all.mean().detach()
    stats["psnr"] = -10.0 * torch.log(color_loss_all.detach().mean()) / torch.log(torch.tensor([10.0], device=tar.device))
    if ssim is not None:
        s = ssim(output["image"].permute(0, 3, 1, 2), tar.permute(0, 3, 1, 2))
        stats["ssim"] = s.detach()
        loss = loss + 0.5 * (1 - s)
    if "rend_dist" in output and iter > 1000:
        distortion = output["rend_dist"].mean()
        stats["distortion"] = distortion.detach()
        loss = loss + distortion * 1000
        acc_map = output["acc_map"].detach()
        normal_error = ((1 - (output["rend_normal"] * output["depth_normal"]).sum(dim=-1)) * acc_map).mean()
        stats["normal"] = normal_error.detach()
        loss = loss + normal_error * 0.2
    return loss, stats


def decoder_layout_torch(parameters, group_centers, K, sh_dim, opacity_shift, scaling_shift, half_cell_size):
    """Plain-torch restatement of Decoder.forward_coarse after the MLP (lightning/network.py:261-278) followed by
    Network.get_offseted_pt (:425-429).  Returns (centers, sh, scaling, rotation, opacity)."""
    parameters = parameters.view(*parameters.shape[:-1], K, -1)
    offset, sh, opacity, scaling, rotation = torch.split(parameters, [3, sh_dim, 1, 2, 4], dim=-1)
    opacity = opacity + opacity_shift
    scaling = scaling + scaling_shift
    offset = torch.sigmoid(offset) * 2 - 1.0
    B = opacity.shape[0]
    sh = sh.view(B, -1, sh_dim // 3, 3)
    opacity = opacity.view(B, -1, 1)
    scaling = scaling.view(B, -1, 2)
    rotation = rotation.view(B, -1, 4)
    offset = offset.view(B, -1, 3)
    centers = group_centers.reshape(1, -1, 3).unsqueeze(-2).expand(B, -1, K, -1).reshape(offset.shape) + offset * half_cell_size
    return centers, sh, scaling, rotation, opacity
