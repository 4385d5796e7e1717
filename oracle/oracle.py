"""numpy/ctypes wrapper of the CPU oracle (``oracle/surfel_oracle.c``).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never from lara_b200/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_uint8, c_uint32, c_void_p
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class _In(ctypes.Structure):
    _fields_ = [("P", c_int), ("D", c_int), ("M", c_int), ("H", c_int), ("W", c_int),
                ("tanfovx", c_float), ("tanfovy", c_float),
                ("bg", c_void_p), ("means3D", c_void_p), ("shs", c_void_p), ("colors_precomp", c_void_p),
                ("opacities", c_void_p), ("scales", c_void_p), ("rotations", c_void_p),
                ("viewmatrix", c_void_p), ("campos", c_void_p)]


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "surfel_oracle.c")
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB)
        lib.oracle_forward.restype = c_void_p
        lib.oracle_forward.argtypes = [POINTER(_In)]
        lib.oracle_destroy.argtypes = [c_void_p]
        lib.oracle_backward.restype = c_int
        lib.oracle_backward.argtypes = [c_void_p] + [c_void_p] * 10
        lib.oracle_ntiles.restype = c_int
        lib.oracle_ntiles.argtypes = [c_void_p]
        lib.oracle_num_rendered.restype = c_uint32
        lib.oracle_num_rendered.argtypes = [c_void_p]
        lib.oracle_threads.restype = c_int
        for name in ("radii", "depths", "transmat", "center", "normal", "rgb", "clamped", "tiles_touched",
                     "point_list", "ranges", "accum", "n_contrib", "out_color", "out_others"):
            fn = getattr(lib, "oracle_" + name)
            fn.restype = c_void_p
            fn.argtypes = [c_void_p]
        _lib = lib
    return _lib


def _f32(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _view(ptr, dtype, shape):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


class OracleRun:
    """One forward of the oracle; keeps the state alive for ``backward``."""

    def __init__(self, *, means3D, opacities, scales, rotations, viewmatrix, campos, bg,
                 image_height, image_width, tanfovx, tanfovy, shs=None, colors_precomp=None, sh_degree=0):
        lib = load()
        self._keep = dict(
            means3D=_f32(means3D), opacities=_f32(opacities).reshape(-1), scales=_f32(scales),
            rotations=_f32(rotations), viewmatrix=_f32(viewmatrix).reshape(-1), campos=_f32(campos),
            bg=_f32(bg), shs=_f32(shs), colors_precomp=_f32(colors_precomp))
        k = self._keep
        self.P = int(k["means3D"].shape[0])
        self.M = int(k["shs"].shape[1]) if k["shs"] is not None else 0
        self.H, self.W = int(image_height), int(image_width)

        def p(a):
            return a.ctypes.data if a is not None else None

        arg = _In(self.P, int(sh_degree), self.M, self.H, self.W, float(tanfovx), float(tanfovy),
                  p(k["bg"]), p(k["means3D"]), p(k["shs"]), p(k["colors_precomp"]), p(k["opacities"]),
                  p(k["scales"]), p(k["rotations"]), p(k["viewmatrix"]), p(k["campos"]))
        self._h = lib.oracle_forward(ctypes.byref(arg))
        if not self._h:
            raise MemoryError("oracle_forward failed")
        h, P, H, W = self._h, self.P, self.H, self.W
        self.ntiles = lib.oracle_ntiles(h)
        self.num_rendered = int(lib.oracle_num_rendered(h))
        R = self.num_rendered
        self.radii = _view(lib.oracle_radii(h), np.int32, (P,))
        self.depths = _view(lib.oracle_depths(h), np.float32, (P,))
        self.transMat = _view(lib.oracle_transmat(h), np.float32, (P, 9))
        self.center = _view(lib.oracle_center(h), np.float32, (P, 2))
        self.normal = _view(lib.oracle_normal(h), np.float32, (P, 3))
        self.rgb = _view(lib.oracle_rgb(h), np.float32, (P, 3))
        self.clamped = _view(lib.oracle_clamped(h), np.uint8, (P, 3))
        self.tiles_touched = _view(lib.oracle_tiles_touched(h), np.uint32, (P,))
        self.point_list = _view(lib.oracle_point_list(h), np.uint32, (R,)) if P else np.zeros((0,), np.uint32)
        self.ranges = _view(lib.oracle_ranges(h), np.uint32, (self.ntiles, 2))
        self.accum = _view(lib.oracle_accum(h), np.float32, (3, H, W))
        self.n_contrib = _view(lib.oracle_n_contrib(h), np.uint32, (2, H, W))
        self.color = _view(lib.oracle_out_color(h), np.float32, (3, H, W))
        self.allmap = _view(lib.oracle_out_others(h), np.float32, (8, H, W))

    def backward(self, grad_color, grad_allmap) -> Dict[str, np.ndarray]:
        lib = load()
        gc, ga = _f32(grad_color), _f32(grad_allmap)
        P, M = self.P, self.M
        out = {
            "means3D": np.zeros((P, 3), np.float32), "means2D": np.zeros((P, 3), np.float32),
            "sh": np.zeros((P, M, 3), np.float32), "colors": np.zeros((P, 3), np.float32),
            "opacities": np.zeros((P, 1), np.float32), "scales": np.zeros((P, 2), np.float32),
            "rotations": np.zeros((P, 4), np.float32), "transMat": np.zeros((P, 9), np.float32),
        }
        rc = lib.oracle_backward(self._h, gc.ctypes.data, ga.ctypes.data,
                                 out["means3D"].ctypes.data, out["means2D"].ctypes.data, out["sh"].ctypes.data,
                                 out["colors"].ctypes.data, out["opacities"].ctypes.data, out["scales"].ctypes.data,
                                 out["rotations"].ctypes.data, out["transMat"].ctypes.data)
        if rc != 0:
            raise MemoryError("oracle_backward failed")
        return out

    def close(self):
        if getattr(self, "_h", None):
            load().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_scene(sc: dict, cam, bg, **kw) -> OracleRun:
    """Convenience: oracle forward for a lara_b200.scene scene + Camera."""
    return OracleRun(means3D=sc["means3D"], opacities=sc["opacities"], scales=sc["scales"],
                     rotations=sc["rotations"], shs=sc.get("shs"), sh_degree=sc.get("sh_degree", 0),
                     viewmatrix=cam.viewmatrix, campos=cam.campos, bg=bg,
                     image_height=cam.image_height, image_width=cam.image_width,
                     tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, **kw)


def threads() -> int:
    return int(load().oracle_threads())
