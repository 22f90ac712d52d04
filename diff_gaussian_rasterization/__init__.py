"""Drop-in module name: `from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(agents/manigaussian_bc/gaussian_renderer/__init__.py:14 of the reference) resolves to the B200-native
implementation when this repository's root is on sys.path (or the package is installed)."""
from manigaussian_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                          rasterize_gaussians, _RasterizeGaussians)


class _CShim:
    """Stands in for the reference's pybind module `_C` (DGR/ext.cpp:14-18)."""
    from manigaussian_b200.rasterizer import rasterize_gaussians_raw as rasterize_gaussians
    from manigaussian_b200.rasterizer import rasterize_gaussians_backward_raw as rasterize_gaussians_backward
    from manigaussian_b200.rasterizer import mark_visible_raw as mark_visible
    rasterize_gaussians = staticmethod(rasterize_gaussians)
    rasterize_gaussians_backward = staticmethod(rasterize_gaussians_backward)
    mark_visible = staticmethod(mark_visible)


_C = _CShim
