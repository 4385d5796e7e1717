"""Host-side mirror of the reference's Python surface for the rasterizer hot path.

Same names, argument meaning and error behaviour as
``third_party/diff-surfel-rasterization/diff_surfel_rasterization/__init__.py``
(GaussianRasterizationSettings :158-170, GaussianRasterizer :172-222,
rasterize_gaussians :21-42, _RasterizeGaussians :44-156), so that
``lightning/renderer_2dgs.py`` runs unchanged.  Underneath it calls the C ABI in
``include/surfel_rasterizer.h`` through ctypes; torch only provides device memory
and the current CUDA stream.

Differences from the reference that a caller can observe (all documented in
DESIGN.md): work is enqueued on torch's *current* stream instead of the legacy
default stream; omitted optional inputs are passed as NULL instead of
``torch.Tensor([]).cuda()`` (which costs a synchronous H2D copy each);
``grad_colors_precomp`` / ``grad_cov3Ds_precomp`` are returned only when the
corresponding input was given (autograd discards them otherwise).
"""
from __future__ import annotations

import os
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# ---------------------------------------------------------------------------------
# small host-side helpers
# ---------------------------------------------------------------------------------

_size_cache: dict = {}
_capacity_hwm: dict = {}          # device index -> largest num_rendered seen (instances)
_pinned_slots: dict = {}          # device index -> pinned uint32 readback slot


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _opt(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The reference passes empty tensors for "not given"; normalise to None."""
    if t is None or t.numel() == 0:
        return None
    return t


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected scalar type Float but found {t.dtype} for {name}")
    return t.contiguous()


def _sizes(lib, P: int, H: int, W: int):
    key = (P, H, W)
    v = _size_cache.get(key)
    if v is None:
        v = _lib.sizes(lib, P, H, W)
        if len(_size_cache) > 64:
            _size_cache.clear()
        _size_cache[key] = v
    return v


def _blob(nbytes: int, device) -> torch.Tensor:
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def _slot(device: torch.device) -> torch.Tensor:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    s = _pinned_slots.get(idx)
    if s is None:
        s = torch.zeros(1, dtype=torch.int32).pin_memory()
        _pinned_slots[idx] = s
    return s


_events: dict = {}


def _event(device: torch.device) -> "torch.cuda.Event":
    """One reusable CUDA event per device (the forward waits on it before returning)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    e = _events.get(idx)
    if e is None:
        # The forward waits on this event for num_rendered.  Blocking (sleeping) wait by default: the GPU boxes
        # run under a CPU quota (16 CPUs for a 1-GPU box although 128 are visible), so a rank that spins takes
        # cycles from the other ranks' host threads; measured at N=1: 1726 vs 1724 views/s (no difference).
        # SRF_SPIN_EVENT_WAIT=1 restores the spinning wait.
        e = torch.cuda.Event(blocking=os.environ.get("SRF_SPIN_EVENT_WAIT", "0") != "1")
        _events[idx] = e
    return e


def _raw_stream(device: torch.device) -> int:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


class _DeviceGuard:
    """`with torch.cuda.device(d)` only when d is not already current (the common case)."""

    def __init__(self, device: torch.device):
        self.idx = device.index
        self.ctx = None

    def __enter__(self):
        if self.idx is not None and self.idx != torch.cuda.current_device():
            self.ctx = torch.cuda.device(self.idx)
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def initial_capacity(P: int, device: torch.device) -> int:
    """Optimistic instance capacity for the binning buffers of a forward call."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    hwm = _capacity_hwm.get(idx, 0)
    return max(int(hwm * 1.25) + 1024, 8 * P, 1 << 16)


def _note_rendered(R: int, device: torch.device) -> None:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if R > _capacity_hwm.get(idx, 0):
        _capacity_hwm[idx] = R


class ForwardState(NamedTuple):
    """What a forward leaves behind for the backward (the reference's three blobs)."""
    geom: torch.Tensor
    tile: torch.Tensor
    image: torch.Tensor
    point_list: torch.Tensor
    capacity: int
    num_rendered: int


def forward_raw(means3D, shs, colors_precomp, opacities, scales, rotations, transMat_precomp,
                raster_settings: GaussianRasterizationSettings, raw_activations: bool = False):
    """Enqueue one forward on the current stream of the tensors' device.

    Returns (color[3,H,W], allmap[8,H,W], radii[P], ForwardState).  Inputs must already be
    normalised (None for absent, fp32, contiguous, CUDA).
    """
    with _DeviceGuard(means3D.device):
        return _forward_raw(means3D, shs, colors_precomp, opacities, scales, rotations, transMat_precomp,
                            raster_settings, raw_activations)


def _forward_raw(means3D, shs, colors_precomp, opacities, scales, rotations, transMat_precomp, raster_settings,
                 raw_activations=False):
    lib = _lib.load()
    H, W = int(raster_settings.image_height), int(raster_settings.image_width)
    P = int(means3D.shape[0])
    device = means3D.device
    M = int(shs.shape[1]) if shs is not None else 0
    geom_b, tile_b, image_b, _ = _sizes(lib, P, H, W)

    # the three state workspaces share one byte blob (sub-blobs stay 256-byte aligned: the sizes are
    # multiples of 256).  The two images stay separate tensors: outputs of an autograd.Function
    # that are views of a common base cannot be modified in place by the caller.
    color = torch.empty((3, H, W), dtype=torch.float32, device=device)
    allmap = torch.empty((8, H, W), dtype=torch.float32, device=device)
    radii = torch.empty((P,), dtype=torch.int32, device=device)
    blob = _blob(geom_b + tile_b + image_b, device)
    geom, tile, image = blob[:geom_b], blob[geom_b:geom_b + tile_b], blob[geom_b + tile_b:]
    sptr = _raw_stream(device)

    if P == 0:
        # reference: zero-filled outputs, nothing rendered (rasterize_points.cu:92-94,105)
        color.zero_(); allmap.zero_()
        tile.zero_(); image.zero_()
        pl = torch.empty((0,), dtype=torch.int32, device=device)
        return color, allmap, radii, ForwardState(geom, tile, image, pl, 0, 0)

    bg = raster_settings.bg
    viewm = raster_settings.viewmatrix
    projm = raster_settings.projmatrix
    campos = raster_settings.campos
    slot = _slot(device)
    _lib.check(lib.srf_forward_preprocess(
        sptr, P, int(raster_settings.sh_degree), M,
        _ptr(means3D), _ptr(shs), _ptr(colors_precomp),
        _ptr(opacities), _ptr(scales), float(raster_settings.scale_modifier),
        _ptr(rotations), _ptr(transMat_precomp),
        _ptr(viewm), _ptr(projm), _ptr(campos),
        float(raster_settings.tanfovx), float(raster_settings.tanfovy), H, W,
        1 if raster_settings.prefiltered else 0,
        radii.data_ptr(), geom.data_ptr(), tile.data_ptr(), slot.data_ptr(), 1 if raw_activations else 0), lib)
    ev = _event(device)
    ev.record()

    # Optimistic capacity: stage 2 is enqueued before num_rendered is known on the host, so
    # the GPU never idles behind the read-back; an overflow (rare) just re-runs stage 2.
    capacity = initial_capacity(P, device)
    while True:
        ent_b, pl_b = _lib.binning_sizes(lib, capacity)
        entries = _blob(ent_b, device)
        point_list = _blob(pl_b, device)
        _lib.check(lib.srf_forward_render(
            sptr, P, H, W, capacity, geom.data_ptr(), tile.data_ptr(),
            entries.data_ptr(), point_list.data_ptr(), image.data_ptr(),
            _ptr(bg), color.data_ptr(), allmap.data_ptr()), lib)
        ev.synchronize()
        R = int(slot.item()) & 0xFFFFFFFF
        if R <= capacity:
            break
        capacity = int(R * 1.25) + 1024
    _note_rendered(R, device)
    if raster_settings.debug:
        torch.cuda.synchronize(device)
    return color, allmap, radii, ForwardState(geom, tile, image, point_list, capacity, R)


def backward_raw(state: ForwardState, radii, means3D, shs, colors_precomp, scales, rotations,
                 transMat_precomp, raster_settings, grad_color, grad_allmap, *,
                 out: Optional[dict] = None, accumulate: bool = False, need_means2D: bool = True,
                 raw_activations: bool = False):
    """Enqueue one backward; returns a dict of gradient tensors.

    ``out`` may supply pre-allocated (possibly strided-into-a-flat-buffer but contiguous)
    tensors for 'means3D','sh','opacities','scales','rotations'; with ``accumulate`` the
    kernel adds into them (view-sharded accumulation).
    """
    with _DeviceGuard(means3D.device):
        return _backward_raw(state, radii, means3D, shs, colors_precomp, scales, rotations, transMat_precomp,
                             raster_settings, grad_color, grad_allmap, out=out, accumulate=accumulate,
                             need_means2D=need_means2D, raw_activations=raw_activations)


def _backward_raw(state, radii, means3D, shs, colors_precomp, scales, rotations, transMat_precomp,
                  raster_settings, grad_color, grad_allmap, *, out=None, accumulate=False, need_means2D=True,
                  raw_activations=False):
    lib = _lib.load()
    H, W = int(raster_settings.image_height), int(raster_settings.image_width)
    P = int(means3D.shape[0])
    device = means3D.device
    M = int(shs.shape[1]) if shs is not None else 0
    out = out or {}

    shapes = {"means3D": (P, 3), "means2D": (P, 3) if need_means2D else None,
              "sh": (P, M, 3) if shs is not None else None,
              "colors_precomp": (P, 3) if colors_precomp is not None else None,
              "opacities": (P, 1), "scales": (P, 2), "rotations": (P, 4),
              "cov3Ds_precomp": (P, 9) if transMat_precomp is not None else None}
    g = {k: out.get(k) for k in shapes}
    missing = [k for k, shp in shapes.items() if shp is not None and g[k] is None]
    if missing:
        # one allocation for all gradients that the caller did not supply (16-byte aligned segments)
        sizes = [(-(-int(torch.Size(shapes[k]).numel()) // 4)) * 4 for k in missing]
        flat = (torch.zeros if accumulate else torch.empty)(sum(sizes), dtype=torch.float32, device=device)
        o = 0
        for k, sz in zip(missing, sizes):
            n = int(torch.Size(shapes[k]).numel())
            g[k] = flat[o:o + n].view(shapes[k])
            o += sz
    if P == 0:
        return g
    _, _, _, scratch_b = _sizes(lib, P, H, W)
    scratch = _blob(scratch_b, device)
    sptr = _raw_stream(device)
    _lib.check(lib.srf_backward(
        sptr, P, int(raster_settings.sh_degree), M, H, W,
        state.capacity, _ptr(raster_settings.bg),
        _ptr(means3D), _ptr(shs), 1 if colors_precomp is not None else 0,
        _ptr(scales), _ptr(rotations), 1 if transMat_precomp is not None else 0,
        _ptr(raster_settings.viewmatrix), _ptr(raster_settings.projmatrix), _ptr(raster_settings.campos),
        float(raster_settings.tanfovx), float(raster_settings.tanfovy), radii.data_ptr(),
        state.geom.data_ptr(), state.tile.data_ptr(), state.point_list.data_ptr(), state.image.data_ptr(),
        grad_color.data_ptr(), grad_allmap.data_ptr(),
        scratch.data_ptr(), 1 if accumulate else 0,
        _ptr(g["means3D"]), _ptr(g["means2D"]), _ptr(g["sh"]), _ptr(g["colors_precomp"]),
        _ptr(g["opacities"]), _ptr(g["scales"]), _ptr(g["rotations"]), _ptr(g["cov3Ds_precomp"]),
        1 if raw_activations else 0), lib)
    if raster_settings.debug:
        torch.cuda.synchronize(device)
    return g


def _dump_snapshot(path: str, items) -> None:
    """CPU copies of the call's tensors, like the reference's cpu_deep_copy_tuple + torch.save."""
    try:
        def cpu(x):
            if isinstance(x, torch.Tensor):
                return x.detach().cpu().clone()
            if isinstance(x, tuple) and hasattr(x, "_fields"):
                return tuple(cpu(v) for v in x)
            return x
        torch.save(tuple(cpu(i) for i in items), path)
    except Exception:
        pass


def _check_settings(rs: GaussianRasterizationSettings, device) -> GaussianRasterizationSettings:
    """Device / dtype checks of the settings tensors, and `.contiguous()` exactly where the reference
    calls it (rasterize_points.cu:113-131): LaRa's MiniCam passes `w2c.transpose(0, 1)`, a
    non-contiguous view, as the view matrix (lightning/utils.py:40)."""
    fixed = {}
    for name in ("bg", "viewmatrix", "projmatrix", "campos"):
        t = getattr(rs, name)
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if t.dtype != torch.float32:
            raise RuntimeError(f"expected scalar type Float but found {t.dtype} for raster_settings.{name}")
        if not t.is_contiguous():
            fixed[name] = t.contiguous()
    return rs._replace(**fixed) if fixed else rs


def _normalise_inputs(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp):
    sh, colors_precomp = _opt(sh), _opt(colors_precomp)
    scales, rotations, cov3Ds_precomp = _opt(scales), _opt(rotations), _opt(cov3Ds_precomp)
    # shape checks of RasterizeGaussiansCUDA (rasterize_points.cu:61-71)
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if scales is not None and (scales.ndim != 2 or scales.shape[1] != 2):
        raise RuntimeError("scales must have dimensions (num_points, 2)")
    if rotations is not None and (rotations.ndim != 2 or rotations.shape[1] != 4):
        raise RuntimeError("rotations must have dimensions (num_points, 4)")
    P = means3D.shape[0]
    means3D = _f32c(means3D, "means3D")
    opacities = _f32c(opacities, "opacity")
    if opacities.numel() != P:
        raise RuntimeError("opacity must have one entry per point")
    if sh is not None:
        sh = _f32c(sh, "sh")
        if sh.ndim != 3 or sh.shape[0] != P or sh.shape[2] != 3:
            raise RuntimeError("sh must have dimensions (num_points, M, 3)")
    if colors_precomp is not None:
        colors_precomp = _f32c(colors_precomp, "colors")
        if colors_precomp.shape != (P, 3):
            raise RuntimeError("colors must have dimensions (num_points, 3)")
    if scales is not None:
        scales = _f32c(scales, "scales")
    if rotations is not None:
        rotations = _f32c(rotations, "rotations")
    if cov3Ds_precomp is not None:
        cov3Ds_precomp = _f32c(cov3Ds_precomp, "transMat_precomp")
        if cov3Ds_precomp.numel() != P * 9:
            raise RuntimeError("transMat_precomp must have dimensions (num_points, 9)")
    if P > 0:
        if (sh is None) == (colors_precomp is None):
            raise RuntimeError("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3Ds_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3Ds_precomp is not None):
            raise RuntimeError("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    return means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp


# ---------------------------------------------------------------------------------
# the reference's public surface
# ---------------------------------------------------------------------------------

def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    RAW_ACTIVATIONS = False     # the subclass below flips this: inputs are raw network outputs

    @classmethod
    def forward(cls, ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        raster_settings = _check_settings(raster_settings, means3D.device)
        (means3D_c, sh_c, colors_c, opac_c, scales_c, rot_c, cov_c) = _normalise_inputs(
            means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp)
        try:
            color, allmap, radii, state = forward_raw(
                means3D_c, sh_c, colors_c, opac_c, scales_c, rot_c, cov_c, raster_settings,
                raw_activations=cls.RAW_ACTIVATIONS)
        except Exception:
            if raster_settings.debug:
                # same debugging aid as the reference (DSR __init__.py:83-90)
                _dump_snapshot("snapshot_fw.dump", (raster_settings, means3D, sh, colors_precomp, opacities,
                                                   scales, rotations, cov3Ds_precomp))
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
            raise
        ctx.raster_settings = raster_settings
        ctx.num_rendered = state.num_rendered
        ctx.capacity = state.capacity
        ctx.present = (sh_c is not None, colors_c is not None, scales_c is not None,
                       rot_c is not None, cov_c is not None)
        ctx.opac_shape = tuple(opacities.shape)
        ctx.raw_activations = cls.RAW_ACTIVATIONS
        dummy = means3D_c.new_empty(0)
        ctx.save_for_backward(
            colors_c if colors_c is not None else dummy, means3D_c,
            scales_c if scales_c is not None else dummy, rot_c if rot_c is not None else dummy,
            cov_c if cov_c is not None else dummy, radii, sh_c if sh_c is not None else dummy,
            state.geom, state.point_list, state.image, state.tile)
        ctx.mark_non_differentiable(radii)
        return color, radii, allmap

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        rs = ctx.raster_settings
        (colors_c, means3D, scales, rotations, cov_c, radii, sh, geom, point_list, image, tile) = ctx.saved_tensors
        has_sh, has_col, has_sc, has_rot, has_cov = ctx.present
        state = ForwardState(geom, tile, image, point_list, ctx.capacity, ctx.num_rendered)
        grad_out_color = _f32c(grad_out_color, "dL_dout_color")
        grad_depth = _f32c(grad_depth, "dL_dout_others")
        try:
            g = backward_raw(state, radii, means3D,
                             sh if has_sh else None, colors_c if has_col else None,
                             scales if has_sc else None, rotations if has_rot else None,
                             cov_c if has_cov else None, rs, grad_out_color, grad_depth,
                             raw_activations=ctx.raw_activations)
        except Exception:
            if rs.debug:
                _dump_snapshot("snapshot_bw.dump", (rs, means3D, radii, sh, scales, rotations, grad_out_color, grad_depth))
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
            raise
        # order of DSR __init__.py:144-154
        if g["opacities"].shape != ctx.opac_shape:
            g["opacities"] = g["opacities"].view(ctx.opac_shape)
        return (g["means3D"], g["means2D"], g["sh"], g["colors_precomp"], g["opacities"],
                g["scales"], g["rotations"], g["cov3Ds_precomp"], None)


class _RasterizeGaussiansRaw(_RasterizeGaussians):
    """Same call, but opacities / scales / rotations are LaRa's raw network outputs (logits,
    log-scales, unnormalised quaternions): the activations of renderer_2dgs.py:183-188 and their
    vjps run inside the preprocess kernels (next-row extension)."""
    RAW_ACTIVATIONS = True


def rasterize_gaussians_raw(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                            cov3Ds_precomp, raster_settings):
    return _RasterizeGaussiansRaw.apply(means3D, means2D, sh, colors_precomp, opacities, scales,
                                        rotations, cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of points in front of the near plane (DSR __init__.py:177-186)."""
        with torch.no_grad():
            lib = _lib.load()
            pos = _f32c(positions, "means3D")
            rs = _check_settings(self.raster_settings, pos.device)
            P = int(pos.shape[0])
            present = torch.zeros((P,), dtype=torch.bool, device=pos.device)
            if P:
                with torch.cuda.device(pos.device):
                    _lib.check(lib.srf_mark_visible(
                        torch.cuda.current_stream(pos.device).cuda_stream, P, pos.data_ptr(),
                        rs.viewmatrix.data_ptr(), rs.projmatrix.data_ptr(), present.data_ptr()), lib)
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales,
                                   rotations, cov3D_precomp, raster_settings)

    def forward_raw_activations(self, means3D, means2D, opacities, scales, rotations, shs=None, colors_precomp=None):
        """Extension (no reference counterpart): like ``forward`` on the scale/rotation path, but
        opacities / scales / rotations are LaRa's raw network outputs; sigmoid / exp / F.normalize
        (renderer_2dgs.py:183-188) and their vjps run inside the preprocess kernels."""
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if scales is None or rotations is None:
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return rasterize_gaussians_raw(means3D, means2D, shs, colors_precomp, opacities, scales,
                                       rotations, None, self.raster_settings)
