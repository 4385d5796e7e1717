"""View-sharded rasterization: the multi-GPU form of the hot path.

Every target view of a scene is an independent forward+backward over the same read-only
Gaussian set; only the parameter gradients sum across views.  One process per GPU
(``torch.distributed``, NCCL over NVLink): rank r renders views ``r, r+G, r+2G, ...``,
the backward kernel *accumulates* each view's parameter gradients straight into one flat
fp32 buffer (segment-major: means3D | sh | opacity | scales | rotations), and a single
``all_reduce(sum)`` of that buffer finishes the step.  Images / aux maps stay on the rank
that owns the view.  The reference has nothing like this (it is single-GPU; LaRa only
data-parallelises over scenes), so there is no reference API to mirror here -- this is an
additional entry point next to the unchanged per-view one.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import rasterizer as R


def shard_views(num_views: int, rank: int, world_size: int) -> List[int]:
    """Indices of the views owned by `rank` (round-robin, so shards differ by <= 1 view)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    return list(range(rank, num_views, world_size))


class GradBuffer:
    """Flat parameter-gradient buffer with per-parameter contiguous views."""

    ORDER = ("means3D", "sh", "opacities", "scales", "rotations")

    def __init__(self, P: int, M: int, device, dtype=torch.float32):
        self.P, self.M = P, M
        sizes = {"means3D": 3 * P, "sh": 3 * M * P, "opacities": P, "scales": 2 * P, "rotations": 4 * P}
        shapes = {"means3D": (P, 3), "sh": (P, M, 3), "opacities": (P, 1), "scales": (P, 2), "rotations": (P, 4)}
        # every segment starts on a 16-byte boundary (the kernel uses vector accesses)
        offs, o = {}, 0
        for k in self.ORDER:
            offs[k] = o
            o += (sizes[k] + 3) // 4 * 4
        self.flat = torch.zeros(o, dtype=dtype, device=device)
        self.views: Dict[str, torch.Tensor] = {
            k: self.flat[offs[k]:offs[k] + sizes[k]].view(*shapes[k]) for k in self.ORDER}

    def zero_(self):
        self.flat.zero_()
        return self

    def all_reduce(self, group=None):
        """The one collective of the view-sharded step: sum of parameter gradients."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self


def render_views(params: Dict[str, torch.Tensor], settings_list: Sequence[R.GaussianRasterizationSettings],
                 upstream: Callable[[int, torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                 grads: Optional[GradBuffer] = None, view_ids: Optional[Sequence[int]] = None,
                 raster_fn=None, streams: int = 1, _pool: Optional[dict] = None):
    """Forward+backward of the given views, accumulating parameter gradients into `grads`.

    params   : dict with contiguous fp32 CUDA tensors means3D [P,3], shs [P,M,3], opacities [P,1],
               scales [P,2], rotations [P,4] (already activated, as LaRa's renderer passes them)
    upstream : callback (view_id, color[3,H,W], allmap[8,H,W]) -> (dL_dcolor, dL_dallmap); this is
               where the caller's loss lives (it runs on the stream the view was rendered on)
    streams  : >1 renders the views round-robin on that many CUDA streams.  Views are independent,
               so the latency-bound per-view kernels (preprocess, tile scan, scatter, per-tile
               sort) and the tail of one view's blend overlap with another view's blend.  Each
               stream accumulates into its own GradBuffer; they are summed into `grads` at the end.
    raster_fn: test hook -- a callable pair replacing (forward_raw, backward_raw)
    Returns (list of (color, allmap, radii) per view, grads).
    """
    P = int(params["means3D"].shape[0])
    M = int(params["shs"].shape[1])
    dev = params["means3D"].device
    if grads is None:
        grads = GradBuffer(P, M, dev)
    fwd, bwd = raster_fn if raster_fn is not None else (R.forward_raw, R.backward_raw)
    outs = []
    ids = list(view_ids) if view_ids is not None else list(range(len(settings_list)))

    def one_view(vid, rs, gbuf):
        if raster_fn is None:
            rs = R._check_settings(rs, dev)      # contiguous fp32 CUDA settings tensors, as the per-view path
        color, allmap, radii, state = fwd(params["means3D"], params["shs"], None, params["opacities"],
                                          params["scales"], params["rotations"], None, rs)
        g_color, g_allmap = upstream(vid, color, allmap)
        bwd(state, radii, params["means3D"], params["shs"], None, params["scales"], params["rotations"], None,
            rs, g_color, g_allmap, out=gbuf.views, accumulate=True, need_means2D=False)
        return color, allmap, radii

    n_streams = max(1, min(int(streams), len(ids)))
    if n_streams == 1 or not params["means3D"].is_cuda:
        for vid, rs in zip(ids, settings_list):
            outs.append(one_view(vid, rs, grads))
        return outs, grads

    pool = _pool if _pool is not None else _STREAM_POOL
    key = (dev.index, n_streams, P, M)
    if key not in pool:
        pool[key] = ([torch.cuda.Stream(device=dev) for _ in range(n_streams)],
                     [GradBuffer(P, M, dev) for _ in range(n_streams - 1)])
    side_streams, side_bufs = pool[key]
    main = torch.cuda.current_stream(dev)
    bufs = [grads] + list(side_bufs)
    for b in side_bufs:
        b.zero_()
    for s in side_streams:
        s.wait_stream(main)
    for k, (vid, rs) in enumerate(zip(ids, settings_list)):
        s = side_streams[k % n_streams]
        with torch.cuda.stream(s):
            o = one_view(vid, rs, bufs[k % n_streams])
        for t in o:
            t.record_stream(main)
        outs.append(o)
    for s in side_streams:
        main.wait_stream(s)
    for b in side_bufs:
        grads.flat.add_(b.flat)
    return outs, grads


_STREAM_POOL: dict = {}
