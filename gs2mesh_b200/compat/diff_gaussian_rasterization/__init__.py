"""Drop-in for the reference's `diff_gaussian_rasterization` Python package (forward path).

Put `gs2mesh_b200/compat` on sys.path ahead of the reference submodule and
third_party/gaussian-splatting/gaussian_renderer/__init__.py:14,36-93 runs unchanged on the
B200-native kernels:

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

Mirrors DGR/diff_gaussian_rasterization/__init__.py:157-220 (settings tuple, argument
validation and messages, empty-tensor sentinels for absent optionals).  Forward only: gs2mesh
renders under torch.no_grad() (gs2mesh_utils/renderer_utils.py:374); the backward pass is out
of scope and asking for gradients raises.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from gs2mesh_b200 import rasterizer as _rast


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


def _absent(t):
    return t is None or (isinstance(t, torch.Tensor) and t.numel() == 0)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
    """Same positional signature as the reference's module-level helper (__init__.py:21-42).
    `means2D` only carries gradients in the reference and is ignored here."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad
                                       for t in (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp)):
        raise RuntimeError("gs2mesh_b200's rasterizer is forward-only; call it under torch.no_grad() "
                           "(as gs2mesh does, renderer_utils.py:374)")
    rs = raster_settings
    flags = _rast.DEFAULT_FLAGS | (_rast._lib.RASTER_DEBUG_SYNC if rs.debug else 0)
    out = _rast.rasterize_forward(
        means3D=means3D, opacities=opacities, viewmatrix=rs.viewmatrix, projmatrix=rs.projmatrix, campos=rs.campos, bg=rs.bg,
        width=rs.image_width, height=rs.image_height, tan_fovx=rs.tanfovx, tan_fovy=rs.tanfovy,
        shs=None if _absent(sh) else sh, colors_precomp=None if _absent(colors_precomp) else colors_precomp,
        scales=None if _absent(scales) else scales, rotations=None if _absent(rotations) else rotations,
        cov3D_precomp=None if _absent(cov3Ds_precomp) else cov3Ds_precomp, sh_degree=rs.sh_degree,
        scale_modifier=rs.scale_modifier, prefiltered=rs.prefiltered, flags=flags, want_depth=False, want_final_T=False)
    return out["color"], out["radii"]


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _rast.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])
        return rasterize_gaussians(
            means3D, means2D,
            empty if shs is None else shs,
            empty if colors_precomp is None else colors_precomp,
            opacities,
            empty if scales is None else scales,
            empty if rotations is None else rotations,
            empty if cov3D_precomp is None else cov3D_precomp,
            self.raster_settings)
