"""Drop-in package name expected by LaRa (``lightning/renderer_2dgs.py:7-10``):

    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer

Everything is implemented in :mod:`lara_b200`; this module only re-exports the
reference's public names so that ``renderer_2dgs.py`` imports it unchanged."""
from lara_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    _RasterizeGaussians,
)
