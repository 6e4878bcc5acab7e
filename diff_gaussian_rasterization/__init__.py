"""Alias under the upstream 3DGS name that BASELINE.json's north_star uses."""
from gaustudio_diff_gaussian_rasterization import *  # noqa: F401,F403
from gaustudio_diff_gaussian_rasterization import (GaussianRasterizationSettings, GaussianRasterizer, _C,  # noqa: F401
                                                   _RasterizeGaussians, rasterize_gaussians)
