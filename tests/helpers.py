"""Shared test helpers (not a test module)."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    m = np.abs(b).max()
    return float(np.abs(a - b).max() / (m if m > 0 else 1.0))


def to_dev(sc, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in sc.items()}


def tile_pixel_mask(ranges, H, W):
    """[H,W] bool: pixel lies in a tile whose range is non-empty (the reference leaves the median
    contributor plane uninitialised for empty tiles)."""
    gx = (W + 15) // 16
    nonempty = (ranges[:, 1] > ranges[:, 0])
    ty = np.arange(H) // 16
    tx = np.arange(W) // 16
    return nonempty[(ty[:, None] * gx + tx[None, :])]


def run_candidate(sc, cam, bg, dev, grads=None, debug=False):
    """Candidate forward (+ backward if grads=(gc,ga)) through the raw API; returns a dict of
    CPU numpy arrays including the internal state."""
    from lara_b200 import rasterizer as R
    from lara_b200 import scene as S
    from lara_b200.debug import unpack_state
    scd = to_dev(sc, dev)
    st = S.settings_for(cam, bg, sc["sh_degree"], dev, R.GaussianRasterizationSettings, debug=debug)
    H, W = cam.image_height, cam.image_width
    P = sc["means3D"].shape[0]
    color, allmap, radii, state = R.forward_raw(scd["means3D"], scd.get("shs"), scd.get("colors_precomp"),
                                                scd["opacities"], scd["scales"], scd["rotations"], None, st)
    torch.cuda.synchronize()
    u = unpack_state(state, P, H, W)
    out = {"color": color, "allmap": allmap, "radii": radii, "num_rendered": state.num_rendered}
    out.update({k: u[k] for k in ("ranges", "point_list", "n_contrib", "accum")})
    if P > 0:
        out.update({k: u[k] for k in ("tiles_touched", "depths", "transMat", "means2D", "rgb", "normal")})
    if grads is not None:
        gc, ga = grads
        g = R.backward_raw(state, radii, scd["means3D"], scd.get("shs"), scd.get("colors_precomp"), scd["scales"],
                           scd["rotations"], None, st, gc.to(dev), ga.to(dev))
        torch.cuda.synchronize()
        out.update({"g_" + k: v for k, v in g.items() if v is not None})
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def scene_from_golden(z):
    sc = {"means3D": torch.from_numpy(z["means3D"]), "scales": torch.from_numpy(z["scales"]),
          "rotations": torch.from_numpy(z["rotations"]), "opacities": torch.from_numpy(z["opacities"]),
          "shs": torch.from_numpy(z["shs"]), "sh_degree": int(z["sh_degree"])}
    from lara_b200.scene import Camera
    cam = Camera(int(z["H"]), int(z["W"]), float(z["tanfovx"]), float(z["tanfovy"]),
                 torch.from_numpy(z["viewmatrix"]), torch.from_numpy(z["projmatrix"]), torch.from_numpy(z["campos"]),
                 torch.eye(4))
    return sc, cam, torch.from_numpy(z["bg"])
