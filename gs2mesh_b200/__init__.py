"""gs2mesh_b200 -- B200-native (sm_100a) implementation of gs2mesh's data-parallel hot path:
the forward Gaussian-splat rasterizer behind every stereo pair and the TSDF voxel integration
that fuses the resulting depth maps.  See DESIGN.md for the scope and INTEGRATION.md for how it
drops into the reference.

Importing the package needs neither a GPU nor the built library; every compute entry point
does (there is no CPU fallback).
"""
from . import camera, scene  # noqa: F401  (host-only helpers)

__all__ = ["camera", "scene", "Renderer", "TSDF", "TSDFVolume", "rasterize_forward"]


def __getattr__(name):  # lazy: torch / CUDA-facing modules
    if name == "Renderer":
        from .renderer import Renderer

        return Renderer
    if name in ("TSDF", "TSDFVolume"):
        from . import tsdf

        return getattr(tsdf, name)
    if name == "rasterize_forward":
        from .rasterizer import rasterize_forward

        return rasterize_forward
    raise AttributeError(name)
