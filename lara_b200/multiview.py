"""Batched multi-view entry: all target views of one scene in one autograd node.

LaRa's step renders the V target views of a scene one after the other (``lightning/network.py:486-495``:
``for j, c2w in enumerate(tar_c2ws): frame = self.gs_render.render_img(cam, rays_d, _centers, ...)``) and
then concatenates the frames (``:525``).  Each view is its own autograd node, so autograd adds V gradient
tensors per parameter and nothing overlaps.  ``render_scene_views`` is the same computation as one
``torch.autograd.Function``:

* forward : per view  K1..K6 (+ fused activations) -> fused epilogue, views round-robin on a small pool
            of CUDA streams (the latency-bound binning kernels and one view's blend tail overlap with
            another view's blend), outputs written into stacked ``[V, C, H, W]`` buffers;
* backward: per view  fused epilogue backward -> K7..K9 *accumulating* the parameter gradients in the
            kernels into one flat buffer per stream (``accumulate=1``), summed once at the end.

The per-view state (tile lists, records, ``n_contrib`` ...) is kept alive between the two phases, like the
reference keeps its three blobs in ``ctx`` (DSR ``__init__.py:97``).  An *additional* entry point: the
per-view drop-in API is unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import rasterizer as R
from .epilogue import epilogue_backward_raw, epilogue_forward_raw
from .sharded import GradBuffer

OUT_KEYS = ("image", "depth", "acc_map", "rend_normal", "depth_normal", "rend_dist")
_STATE_FIELDS = 4          # geom, tile, image, point_list
_POOL: dict = {}


def _streams(dev, n):
    key = (dev.index, n)
    if key not in _POOL:
        _POOL[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
    return _POOL[key]


class _RenderSceneViews(torch.autograd.Function):
    @staticmethod
    def forward(ctx, centers, shs, opacity, scales, rotations, rays, settings_list, depth_ratio, raw_activations, n_streams):
        dev = centers.device
        V = len(settings_list)
        P = int(centers.shape[0])
        centers_c, shs_c, _, opac_c, scales_c, rot_c, _ = R._normalise_inputs(
            centers, shs, None, opacity, scales, rotations, None)
        settings_list = [R._check_settings(rs, dev) for rs in settings_list]
        H, W = int(settings_list[0].image_height), int(settings_list[0].image_width)
        for rs in settings_list:
            if (int(rs.image_height), int(rs.image_width)) != (H, W):
                raise RuntimeError("render_scene_views: all views of a call must share one image size")
        rays_c = None
        if rays is not None:
            rays_c = rays.contiguous()
            if rays_c.dtype != torch.float32 or tuple(rays_c.shape) != (V, H, W, 6):
                raise RuntimeError(f"rays must be float32 [V,H,W,6] = {(V, H, W, 6)}, got {tuple(rays_c.shape)}")

        def new(*shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)
        image, depth, acc = new(V, 3, H, W), new(V, 1, H, W), new(V, H, W)
        rn, dn, dist = new(V, 3, H, W), new(V, 3, H, W), new(V, H, W)
        radii = torch.empty((V, P), dtype=torch.int32, device=dev)

        n = max(1, min(int(n_streams), V))
        main = torch.cuda.current_stream(dev)
        side = _streams(dev, n) if n > 1 else [main]
        for s in side:
            if s is not main:
                s.wait_stream(main)
        saved: List[torch.Tensor] = []
        meta = []
        for v, rs in enumerate(settings_list):
            with torch.cuda.stream(side[v % n]):
                color, allmap, rad, st = R.forward_raw(centers_c, shs_c, None, opac_c, scales_c, rot_c, None, rs,
                                                       raw_activations=raw_activations)
                radii[v].copy_(rad)
                if rays_c is not None:
                    epilogue_forward_raw(color, allmap, rays_c[v], rs.viewmatrix, depth_ratio,
                                         image[v], depth[v], acc[v], rn[v], dn[v], dist[v])
                else:
                    image[v].copy_(color.clamp(0, 1))
            saved += [color, allmap, rad, st.geom, st.tile, st.image, st.point_list]
            meta.append((st.capacity, st.num_rendered))
        for s in side:
            if s is not main:
                main.wait_stream(s)

        ctx.save_for_backward(centers_c, shs_c, scales_c, rot_c, rays_c if rays_c is not None else centers_c.new_empty(0), *saved)
        ctx.settings_list = settings_list
        ctx.meta = meta
        ctx.has_rays = rays_c is not None
        ctx.depth_ratio = float(depth_ratio)
        ctx.raw_activations = bool(raw_activations)
        ctx.n_streams = n
        ctx.opac_shape = tuple(opacity.shape)
        ctx.mark_non_differentiable(radii)
        return image, depth, acc, rn, dn, dist, radii

    @staticmethod
    def backward(ctx, g_image, g_depth, g_acc, g_rn, g_dn, g_dist, _g_radii):
        centers, shs, scales, rot, rays, *saved = ctx.saved_tensors
        dev = centers.device
        settings_list = ctx.settings_list
        P, M = int(centers.shape[0]), int(shs.shape[1])

        def c(t):
            return None if t is None else t.contiguous()
        g_image, g_depth, g_acc, g_rn, g_dn, g_dist = c(g_image), c(g_depth), c(g_acc), c(g_rn), c(g_dn), c(g_dist)

        def at(t, v):
            return None if t is None else t[v]

        n = ctx.n_streams
        main = torch.cuda.current_stream(dev)
        side = _streams(dev, n) if n > 1 else [main]
        bufs = [GradBuffer(P, M, dev) for _ in range(n)]          # zero-filled on `main`
        for s in side:
            if s is not main:
                s.wait_stream(main)
        per = 3 + _STATE_FIELDS
        for v, rs in enumerate(settings_list):
            color, allmap, rad, geom, tile, image_state, point_list = saved[v * per:(v + 1) * per]
            cap, num = ctx.meta[v]
            state = R.ForwardState(geom, tile, image_state, point_list, cap, num)
            with torch.cuda.stream(side[v % n]):
                if ctx.has_rays:
                    d_color, d_allmap = epilogue_backward_raw(
                        color, allmap, rays[v], rs.viewmatrix, ctx.depth_ratio, at(g_image, v), at(g_depth, v),
                        at(g_acc, v), at(g_rn, v), at(g_dn, v), at(g_dist, v))
                else:
                    gi = g_image[v] if g_image is not None else torch.zeros_like(color)
                    d_color = gi * ((color >= 0) & (color <= 1))      # vjp of clamp(0, 1)
                    d_allmap = torch.zeros_like(allmap)
                R.backward_raw(state, rad, centers, shs, None, scales, rot, None, rs, d_color, d_allmap,
                               out=bufs[v % n].views, accumulate=True, need_means2D=False,
                               raw_activations=ctx.raw_activations)
        for s in side:
            if s is not main:
                main.wait_stream(s)
        total = bufs[0]
        for b in bufs[1:]:
            total.flat.add_(b.flat)
        g = total.views
        return (g["means3D"], g["sh"], g["opacities"].view(ctx.opac_shape), g["scales"], g["rotations"],
                None, None, None, None, None)


def render_scene_views(centers, shs, opacity, scales, rotations, settings_list: Sequence[R.GaussianRasterizationSettings],
                       rays: Optional[torch.Tensor] = None, depth_ratio: float = 0.0, raw_activations: bool = False,
                       streams: int = 3, prex: str = "") -> Dict[str, torch.Tensor]:
    """All views of one scene: returns stacked per-view results, ``{key+prex: [V, H, W, C] or [V, H, W]}`` with
    the keys, channel-last layout and values of ``Renderer.render_img``'s dict (renderer_2dgs.py:256-268) plus
    ``radii+prex`` [V, P]; ``out[k][j]`` is what the j-th ``render_img`` call of network.py:486-495 returns.
    ``rays=None``: only ``image`` as [V,3,H,W] (render_img's early return is the clamped planar image)."""
    if len(settings_list) == 0:
        raise ValueError("render_scene_views needs at least one view")
    image, depth, acc, rn, dn, dist, radii = _RenderSceneViews.apply(
        centers, shs, opacity, scales, rotations, rays, list(settings_list), float(depth_ratio), bool(raw_activations),
        int(streams))
    if rays is None:          # render_img's early return hands back the clamped [3,H,W] image, not a dict
        return {f"image{prex}": image, f"radii{prex}": radii}
    return {
        f"image{prex}": image.permute(0, 2, 3, 1), f"depth{prex}": depth.permute(0, 2, 3, 1), f"acc_map{prex}": acc,
        f"rend_normal{prex}": rn.permute(0, 2, 3, 1), f"depth_normal{prex}": dn.permute(0, 2, 3, 1),
        f"rend_dist{prex}": dist, f"radii{prex}": radii,
    }


def concat_views(out: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The layout LaRa's step hands to the loss: ``torch.cat([frame[k] for frame in views], dim=1)``
    (network.py:525), i.e. the views side by side along the width."""
    res = {}
    for k, t in out.items():
        if k.startswith("radii"):
            continue
        if t.ndim == 4:      # [V,H,W,C] -> [H, V*W, C]
            V, H, W, C = t.shape
            res[k] = t.permute(1, 0, 2, 3).reshape(H, V * W, C)
        else:                # [V,H,W] -> [H, V*W]
            V, H, W = t.shape
            res[k] = t.permute(1, 0, 2).reshape(H, V * W)
    return res
