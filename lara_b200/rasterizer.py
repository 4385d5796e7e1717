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
import threading
import warnings
from typing import List, NamedTuple, Optional, Sequence

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
_capacity_hwm: dict = {}          # device index -> largest num_rendered seen (instances per view)
_lock = threading.Lock()          # autograd calls backward() from another host thread


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


def _dev_index(device: torch.device) -> int:
    return device.index if device.index is not None else torch.cuda.current_device()


def _raw_stream(device: torch.device) -> int:
    return torch._C._cuda_getCurrentRawStream(_dev_index(device))


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
    """Optimistic per-view instance capacity for the binning buffers of a forward call."""
    hwm = _capacity_hwm.get(_dev_index(device), 0)
    return max(int(hwm * 1.25) + 1024, 8 * P, 1 << 16)


def _note_rendered(R: int, device: torch.device) -> None:
    idx = _dev_index(device)
    with _lock:
        if R > _capacity_hwm.get(idx, 0):
            _capacity_hwm[idx] = R


# ---------------------------------------------------------------------------------
# num_rendered read-back.  The reference blocks the host on a cudaMemcpy in the middle of every
# forward (rasterizer_impl.cu:282).  Here the count travels to a pinned slot asynchronously and is
# only *needed* by the host to verify the optimistic instance capacity, so the wait is deferred:
# a forward returns without touching the event; the count is resolved the first time somebody asks
# for it (ForwardState.num_rendered / ctx.num_rendered), when the backward of that forward starts, or
# -- without blocking -- at the next forward on the device once its event has completed.  If the
# capacity turns out to have been too small (first call of a much larger scene), stage 2 is re-run in
# place with larger buffers and a RuntimeWarning says that results consumed in between were stale.
# raster_settings.debug=True (the reference's "synchronise and check" switch) or
# SRF_SYNC_NUM_RENDERED=1 restore the eager wait inside the forward.
# ---------------------------------------------------------------------------------

_SLOT_WORDS = 32
_readback_pool: dict = {}         # device index -> list of free (pinned int32[_SLOT_WORDS], event)
_pending: dict = {}               # device index -> list of unresolved _Pending


def _take_readback(device: torch.device, words: int):
    idx = _dev_index(device)
    with _lock:
        free = _readback_pool.setdefault(idx, [])
        for k, (slot, ev) in enumerate(free):
            if slot.numel() >= words:
                return free.pop(k)
    slot = torch.zeros(max(words, _SLOT_WORDS), dtype=torch.int32).pin_memory()
    # Spinning wait (cudaEventSynchronize without cudaEventBlockingSync).  Whoever resolves a count waits at most for the
    # tile scan of that forward, and the thread has nothing else to do meanwhile.  A sleeping wait costs a wake-up: with 8
    # ranks on one box some rank wakes ~0.2 ms late every step, its GPU runs dry and all ranks wait for it at the
    # all-reduce (measured: 14 979 vs 15 460 views/s on 8 B200s).  SRF_BLOCKING_EVENT_WAIT=1 restores the sleeping wait.
    ev = torch.cuda.Event(blocking=os.environ.get("SRF_BLOCKING_EVENT_WAIT", "0") == "1")
    ev.record()          # materialises the cudaEvent_t so that the C entry point can record it between its two stages
    return slot, ev


def _give_readback(device: torch.device, rb) -> None:
    with _lock:
        _readback_pool.setdefault(_dev_index(device), []).append(rb)


def _eager_sync(debug: bool) -> bool:
    return bool(debug) or os.environ.get("SRF_SYNC_NUM_RENDERED", "0") == "1"


class _Pending:
    """An enqueued forward whose instance counts have not been read back yet."""

    def __init__(self, device, rb, nviews: int, capacity: int, rerun):
        self.device, self.rb, self.V, self.capacity, self.rerun = device, rb, nviews, capacity, rerun
        self.counts: Optional[List[int]] = None
        self.lock = threading.Lock()

    def ready(self) -> bool:
        return self.counts is not None or self.rb[1].query()

    def resolve(self, stale_ok: bool = False) -> List[int]:
        with self.lock:
            if self.counts is not None:
                return self.counts
            slot, ev = self.rb
            ev.synchronize()
            counts = [int(x) & 0xFFFFFFFF for x in slot[:self.V].tolist()]
            _give_readback(self.device, self.rb)
            self.rb = None
            R = max(counts) if counts else 0
            _note_rendered(R, self.device)
            if R > self.capacity:
                self.rerun(int(R * 1.25) + 1024)
                if not stale_ok:
                    warnings.warn(
                        f"surfel rasterizer: {R} instances exceeded the optimistic capacity {self.capacity}; the forward "
                        "was re-run in place, but work enqueued between the two runs read incomplete images. Set "
                        "raster_settings.debug=True or SRF_SYNC_NUM_RENDERED=1 to verify the capacity inside every forward.",
                        RuntimeWarning, stacklevel=3)
            self.rerun = None
            self.counts = counts
        with _lock:
            lst = _pending.get(_dev_index(self.device))
            if lst and self in lst:
                lst.remove(self)
        return counts


def _drain_ready(device: torch.device) -> None:
    """Resolve, without blocking, the pending forwards of this device whose read-back has landed."""
    with _lock:
        lst = list(_pending.get(_dev_index(device), ()))
    for p in lst:
        if p.ready():
            p.resolve()


class LazyCount:
    """int-like view of a forward's num_rendered that waits for the read-back only when asked."""

    def __init__(self, pending: "_Pending", view: int = 0):
        self._p, self._v = pending, view

    def __int__(self):
        return self._p.resolve()[self._v]

    __index__ = __int__

    def __eq__(self, o):
        return int(self) == o

    def __hash__(self):
        return hash(int(self))

    def __repr__(self):
        return str(int(self))


class ForwardState:
    """What a forward leaves behind for the backward (the reference's three blobs).  For the batched
    entry the workspaces hold V per-view workspaces back to back."""
    __slots__ = ("geom", "tile", "image", "point_list", "capacity", "nviews", "_pending", "_counts")

    def __init__(self, geom, tile, image, point_list, capacity, num_rendered=0, nviews: int = 1, pending=None):
        self.geom, self.tile, self.image, self.point_list = geom, tile, image, point_list
        self.capacity, self.nviews = capacity, nviews
        self._pending = pending
        self._counts = None if pending is not None else ([int(num_rendered)] * nviews
                                                         if not isinstance(num_rendered, (list, tuple)) else list(num_rendered))

    def resolve(self) -> List[int]:
        if self._counts is None:
            self._counts = self._pending.resolve()
            self._pending = None
        return self._counts

    @property
    def num_rendered(self):
        c = self.resolve()
        return c[0] if self.nviews == 1 else list(c)


def _launch_forward(device, lib, nviews, P, capacity, debug, call_pre, call_render, state_blobs, alloc_binning,
                    call_both=None):
    """Common enqueue sequence of the per-view and the batched forward:
    stage 1 (preprocess + tile scan + async count read-back) and stage 2 (binning + blend) with an optimistic
    capacity, then an event for whoever eventually asks for the instance counts."""
    _drain_ready(device)
    rb = _take_readback(device, nviews)
    slot, ev = rb
    geom, tile, image = state_blobs
    entries, point_list = alloc_binning(capacity)
    # the event sits between stage 1 (which produces the counts) and stage 2: whoever resolves the counts later waits
    # for the tile scan, never for the blend
    if call_both is not None:
        call_both(slot.data_ptr(), ev.cuda_event, capacity, entries, point_list)
    else:
        call_pre(slot.data_ptr())
        ev.record()
        call_render(capacity, entries, point_list)
    state = ForwardState(geom, tile, image, point_list, capacity, nviews=nviews)

    def rerun(new_capacity: int):
        with _DeviceGuard(device):
            e2, pl2 = alloc_binning(new_capacity)
            call_render(new_capacity, e2, pl2)
            state.point_list, state.capacity = pl2, new_capacity

    pend = _Pending(device, rb, nviews, capacity, rerun)
    state._pending, state._counts = pend, None
    if _eager_sync(debug):
        pend.resolve(stale_ok=True)      # nothing has consumed the outputs yet
        state.resolve()
    else:
        with _lock:
            _pending.setdefault(_dev_index(device), []).append(pend)
    return state


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

    def call_both(slot_ptr, ev_handle, capacity, entries, point_list):
        _lib.check(lib.srf_forward(
            sptr, P, int(raster_settings.sh_degree), M,
            _ptr(means3D), _ptr(shs), _ptr(colors_precomp),
            _ptr(opacities), _ptr(scales), float(raster_settings.scale_modifier),
            _ptr(rotations), _ptr(transMat_precomp),
            _ptr(viewm), _ptr(projm), _ptr(campos),
            float(raster_settings.tanfovx), float(raster_settings.tanfovy), H, W,
            1 if raster_settings.prefiltered else 0, _ptr(bg), capacity,
            radii.data_ptr(), geom.data_ptr(), tile.data_ptr(), entries.data_ptr(), point_list.data_ptr(), image.data_ptr(),
            color.data_ptr(), allmap.data_ptr(), slot_ptr, ev_handle, 1 if raw_activations else 0), lib)

    def alloc_binning(capacity):
        ent_b, pl_b = _lib.binning_sizes(lib, capacity)
        return _blob(ent_b, device), _blob(pl_b, device)

    def call_render(capacity, entries, point_list):
        # (re-)enqueued on whatever stream is current: a deferred re-run happens on the caller's stream
        _lib.check(lib.srf_forward_render(
            _raw_stream(device), P, H, W, capacity, geom.data_ptr(), tile.data_ptr(),
            entries.data_ptr(), point_list.data_ptr(), image.data_ptr(),
            _ptr(bg), color.data_ptr(), allmap.data_ptr()), lib)

    # Optimistic capacity: stage 2 is enqueued before num_rendered is known on the host, so
    # the GPU never idles behind the read-back; an overflow (rare) just re-runs stage 2.
    state = _launch_forward(device, lib, 1, P, initial_capacity(P, device), raster_settings.debug,
                            None, call_render, (geom, tile, image), alloc_binning, call_both=call_both)
    if raster_settings.debug:
        torch.cuda.synchronize(device)
    return color, allmap, radii, state


def pack_cameras(settings_list: Sequence[GaussianRasterizationSettings], device) -> torch.Tensor:
    """[V, SRF_CAM_FLOATS] device tensor of per-view camera records (viewmatrix | campos | bg | pad) built
    from per-view settings with ONE cat kernel (the tensors are already on the device, as LaRa's MiniCam
    produces them, lightning/utils.py:33-48)."""
    parts = []
    pad = torch.zeros(_lib.CAM_FLOATS - 22, dtype=torch.float32, device=device)
    for rs in settings_list:
        parts += [rs.viewmatrix.reshape(16), rs.campos.reshape(3), rs.bg.reshape(3), pad]
    return torch.cat(parts).view(len(settings_list), _lib.CAM_FLOATS)


def forward_views_raw(means3D, shs, colors_precomp, opacities, scales, rotations, transMat_precomp,
                      cams: torch.Tensor, tanfovx: float, tanfovy: float, H: int, W: int, sh_degree: int,
                      prefiltered: bool = False, debug: bool = False, raw_activations: bool = False):
    """All V views of one Gaussian set in ONE launch set (srf_views_*): every kernel carries a view
    dimension.  `cams` is the [V, 24] device tensor of camera records (pack_cameras).
    Returns (color [V,3,H,W], allmap [V,8,H,W], radii [V,P], ForwardState with V back-to-back workspaces)."""
    with _DeviceGuard(means3D.device):
        lib = _lib.load()
        device = means3D.device
        V, P = int(cams.shape[0]), int(means3D.shape[0])
        M = int(shs.shape[1]) if shs is not None else 0
        if cams.dtype != torch.float32 or not cams.is_cuda or cams.shape[1] != _lib.CAM_FLOATS or not cams.is_contiguous():
            raise RuntimeError(f"cams must be a contiguous float32 CUDA tensor [V,{_lib.CAM_FLOATS}]")
        color = torch.empty((V, 3, H, W), dtype=torch.float32, device=device)
        allmap = torch.empty((V, 8, H, W), dtype=torch.float32, device=device)
        radii = torch.empty((V, P), dtype=torch.int32, device=device)
        geom_b, tile_b, image_b, _, _, _ = _lib.views_sizes(lib, V, P, H, W, 0)
        blob = _blob(geom_b + tile_b + image_b, device)
        geom, tile, image = blob[:geom_b], blob[geom_b:geom_b + tile_b], blob[geom_b + tile_b:]
        if P == 0:
            color.zero_(); allmap.zero_(); tile.zero_(); image.zero_()
            pl = torch.empty((0,), dtype=torch.int32, device=device)
            return color, allmap, radii, ForwardState(geom, tile, image, pl, 0, 0, nviews=V)
        sptr = _raw_stream(device)

        def call_pre(slot_ptr):
            _lib.check(lib.srf_views_forward_preprocess(
                sptr, V, P, int(sh_degree), M, _ptr(means3D), _ptr(shs), _ptr(colors_precomp), _ptr(opacities),
                _ptr(scales), _ptr(rotations), _ptr(transMat_precomp), cams.data_ptr(), float(tanfovx), float(tanfovy),
                H, W, 1 if prefiltered else 0, radii.data_ptr(), geom.data_ptr(), tile.data_ptr(), slot_ptr,
                1 if raw_activations else 0), lib)

        def alloc_binning(capacity):
            _, _, _, ent_b, pl_b, _ = _lib.views_sizes(lib, V, P, H, W, capacity)
            return _blob(ent_b, device), _blob(pl_b, device)

        def call_render(capacity, entries, point_list):
            _lib.check(lib.srf_views_forward_render(
                _raw_stream(device), V, P, H, W, capacity, geom.data_ptr(), tile.data_ptr(), entries.data_ptr(),
                point_list.data_ptr(), image.data_ptr(), cams.data_ptr(), color.data_ptr(), allmap.data_ptr()), lib)

        state = _launch_forward(device, lib, V, P, initial_capacity(P, device), debug, call_pre, call_render,
                                (geom, tile, image), alloc_binning)
        if debug:
            torch.cuda.synchronize(device)
        return color, allmap, radii, state


def _grad_outputs(P, M, device, shs, colors_precomp, transMat_precomp, out, accumulate, need_means2D):
    """Gradient tensors the caller did not supply: one small allocation each (cheaper on the host than slicing views out
    of a flat buffer; the caching allocator hands out 512-byte aligned blocks, so the kernel's vector stores are fine)."""
    new = torch.zeros if accumulate else torch.empty
    g = {"means3D": out.get("means3D"), "means2D": out.get("means2D"), "sh": out.get("sh"),
         "colors_precomp": out.get("colors_precomp"), "opacities": out.get("opacities"), "scales": out.get("scales"),
         "rotations": out.get("rotations"), "cov3Ds_precomp": out.get("cov3Ds_precomp")}
    if g["means3D"] is None:
        g["means3D"] = new((P, 3), dtype=torch.float32, device=device)
    if need_means2D and g["means2D"] is None:
        g["means2D"] = new((P, 3), dtype=torch.float32, device=device)
    if shs is not None and g["sh"] is None:
        g["sh"] = new((P, M, 3), dtype=torch.float32, device=device)
    if colors_precomp is not None and g["colors_precomp"] is None:
        g["colors_precomp"] = new((P, 3), dtype=torch.float32, device=device)
    if g["opacities"] is None:
        g["opacities"] = new((P, 1), dtype=torch.float32, device=device)
    if g["scales"] is None:
        g["scales"] = new((P, 2), dtype=torch.float32, device=device)
    if g["rotations"] is None:
        g["rotations"] = new((P, 4), dtype=torch.float32, device=device)
    if transMat_precomp is not None and g["cov3Ds_precomp"] is None:
        g["cov3Ds_precomp"] = new((P, 9), dtype=torch.float32, device=device)
    if not need_means2D:
        g["means2D"] = out.get("means2D")
    if shs is None:
        g["sh"] = None
    return g


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
    g = _grad_outputs(P, M, device, shs, colors_precomp, transMat_precomp, out or {}, accumulate, need_means2D)
    if P == 0:
        return g
    state.resolve()                 # capacity verified (re-run if it overflowed) before the lists are walked again
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


def backward_views_raw(state: ForwardState, radii, means3D, shs, colors_precomp, scales, rotations, transMat_precomp,
                       cams: torch.Tensor, tanfovx: float, tanfovy: float, H: int, W: int, sh_degree: int,
                       grad_color, grad_allmap, *, out: Optional[dict] = None, accumulate: bool = False,
                       need_means2D: bool = False, raw_activations: bool = False, debug: bool = False):
    """Backward of forward_views_raw: grad_color [V,3,H,W], grad_allmap [V,8,H,W]; the per-Gaussian backward
    sums the V views in registers and writes (accumulate: adds into) every parameter-gradient row once."""
    with _DeviceGuard(means3D.device):
        lib = _lib.load()
        device = means3D.device
        V, P = int(cams.shape[0]), int(means3D.shape[0])
        M = int(shs.shape[1]) if shs is not None else 0
        g = _grad_outputs(P, M, device, shs, colors_precomp, transMat_precomp, out or {}, accumulate, need_means2D)
        if P == 0:
            return g
        state.resolve()
        scratch_b = _lib.views_sizes(lib, V, P, H, W, 0)[5]
        scratch = _blob(scratch_b, device)
        _lib.check(lib.srf_views_backward(
            _raw_stream(device), V, P, int(sh_degree), M, H, W, state.capacity, cams.data_ptr(),
            _ptr(means3D), _ptr(shs), 1 if colors_precomp is not None else 0,
            _ptr(scales), _ptr(rotations), 1 if transMat_precomp is not None else 0,
            float(tanfovx), float(tanfovy), radii.data_ptr(),
            state.geom.data_ptr(), state.tile.data_ptr(), state.point_list.data_ptr(), state.image.data_ptr(),
            grad_color.data_ptr(), grad_allmap.data_ptr(), scratch.data_ptr(), 1 if accumulate else 0,
            _ptr(g["means3D"]), _ptr(g["means2D"]), _ptr(g["sh"]), _ptr(g["colors_precomp"]),
            _ptr(g["opacities"]), _ptr(g["scales"]), _ptr(g["rotations"]), _ptr(g["cov3Ds_precomp"]),
            1 if raw_activations else 0), lib)
        if debug:
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
        ctx.state = state
        # an int for an eager forward, else an int-like object that waits for the read-back when asked
        ctx.num_rendered = state.num_rendered if state._pending is None else LazyCount(state._pending)
        ctx.present = (sh_c is not None, colors_c is not None, scales_c is not None,
                       rot_c is not None, cov_c is not None)
        ctx.opac_shape = tuple(opacities.shape)
        ctx.raw_activations = cls.RAW_ACTIVATIONS
        dummy = means3D_c.new_empty(0)
        ctx.save_for_backward(
            colors_c if colors_c is not None else dummy, means3D_c,
            scales_c if scales_c is not None else dummy, rot_c if rot_c is not None else dummy,
            cov_c if cov_c is not None else dummy, radii, sh_c if sh_c is not None else dummy,
            state.geom, state.image, state.tile)
        ctx.mark_non_differentiable(radii)
        return color, radii, allmap

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        rs = ctx.raster_settings
        (colors_c, means3D, scales, rotations, cov_c, radii, sh, geom, image, tile) = ctx.saved_tensors
        has_sh, has_col, has_sc, has_rot, has_cov = ctx.present
        state = ctx.state     # holds the index list (replaced if a deferred capacity check re-ran stage 2)
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
