"""B200-native differentiable Gaussian-splatting rasterizer for ManiGaussian's rendering hot path.

Public surface = the reference's (`diff_gaussian_rasterization`): GaussianRasterizationSettings,
GaussianRasterizer, rasterize_gaussians.  Importing this package never touches oracle/ or any CPU fallback.
"""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians,  # noqa: F401
                         rasterize_gaussians_raw, rasterize_gaussians_backward_raw, mark_visible_raw)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
