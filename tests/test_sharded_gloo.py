"""world_size-2 gloo test of the view-sharded step's host logic: each rank accumulates its
views' parameter gradients and one all-reduce makes every rank hold the sum over all views.
The rasterizer itself is replaced by a deterministic CPU stand-in through the raster_fn hook
(the CUDA path has no CPU fallback); the sharding / accumulation / collective code is real."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _fake_raster():
    from lara_b200.rasterizer import ForwardState

    def fwd(means3D, shs, colors, opac, scales, rot, cov, rs):
        v = float(rs.tanfovx)          # the view id is smuggled in through the settings
        color = torch.full((3, 4, 4), v)
        allmap = torch.full((8, 4, 4), v)
        return color, allmap, torch.ones(means3D.shape[0], dtype=torch.int32), ForwardState(None, None, None, None, 0, int(v))

    def bwd(state, radii, means3D, shs, colors, scales, rot, cov, rs, g_color, g_allmap, *, out, accumulate, need_means2D):
        assert accumulate and not need_means2D
        w = float(g_color.sum())       # depends on the view through the upstream callback
        out["means3D"] += w
        out["sh"] += 2 * w
        out["opacities"] += 3 * w
        out["scales"] += 4 * w
        out["rotations"] += 5 * w
    return fwd, bwd


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lara_b200 import sharded
        from lara_b200.rasterizer import GaussianRasterizationSettings as GS
        P, M, V = 6, 4, 5
        params = {"means3D": torch.zeros(P, 3), "shs": torch.zeros(P, M, 3), "opacities": torch.zeros(P, 1),
                  "scales": torch.zeros(P, 2), "rotations": torch.zeros(P, 4)}
        ids = sharded.shard_views(V, rank, world)
        z = torch.zeros(4, 4)
        sets = [GS(4, 4, float(i), 1.0, torch.zeros(3), 1.0, z, z, 1, torch.zeros(3), False, False) for i in ids]
        seen = []

        def upstream(vid, color, allmap):
            seen.append(vid)
            assert float(color[0, 0, 0]) == float(vid)
            return torch.full((3, 4, 4), float(vid + 1) / 48.0), torch.zeros(8, 4, 4)

        grads = sharded.GradBuffer(P, M, "cpu")
        outs, grads = sharded.render_views(params, sets, upstream, grads=grads, view_ids=ids, raster_fn=_fake_raster())
        assert seen == ids and len(outs) == len(ids)
        grads.all_reduce()
        total = sum(v + 1 for v in range(V))     # every view contributes (vid+1) once, on exactly one rank
        ok = (torch.allclose(grads.views["means3D"], torch.full((P, 3), float(total))) and
              torch.allclose(grads.views["sh"], torch.full((P, M, 3), 2.0 * total)) and
              torch.allclose(grads.views["rotations"], torch.full((P, 4), 5.0 * total)))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_view_sharded_step_two_ranks_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}
