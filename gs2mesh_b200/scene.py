"""Seeded synthetic scenes + camera rings for tests and bench.py.

There is no dataset or checkpoint in this environment, so the BASELINE.json
configurations are realised as synthetic Gaussian clouds in exactly the tensor
layouts `GaussianModel` exposes after its activations
(third_party/gaussian-splatting/scene/gaussian_model.py:95-115):
    xyz [P,3], features [P,16,3], opacity [P,1] (post-sigmoid),
    scaling [P,3] (post-exp), rotation [P,4] (normalised, w first).
Cameras are produced as the reference's camera dicts (renderer_utils.py:178-206)
through gs2mesh_b200.camera, so they flow through the same pose code as COLMAP poses.

Deviation from SURVEY 8(d), stated: the object radius is 0.8 (not 1.0) and camera
ring radius 2.4 so that the reference's DEFAULT TSDF lattice (voxel_length = 2/512,
argument_utils.py:86; dense equivalent = 512^3 over [-1,1]^3) contains the
surface; everything else follows the survey's generator.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import camera as cam

OBJECT_RADIUS = 0.8
CAMERA_RADIUS = 2.4


@dataclass
class GaussianCloud:
    """Post-activation Gaussian parameters, float32 numpy, reference layouts."""

    xyz: np.ndarray  # [P,3]
    features: np.ndarray  # [P,16,3]
    opacity: np.ndarray  # [P,1]
    scaling: np.ndarray  # [P,3]
    rotation: np.ndarray  # [P,4]
    sh_degree: int = 3

    @property
    def num_points(self) -> int:
        return int(self.xyz.shape[0])


def _surface_radius(dirs: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """Radius-1 sphere blended with 3 low-frequency bumps -> a closed, non-trivial surface."""
    r = np.ones(dirs.shape[0])
    for _ in range(3):
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        amp = rng.uniform(0.05, 0.12)
        freq = rng.integers(2, 4)
        r += amp * np.cos(freq * np.arccos(np.clip(dirs @ axis, -1.0, 1.0)))
    return r / r.max()


def make_gaussians(num_points: int, seed: int = 0, sh_degree: int = 3) -> GaussianCloud:
    rng = np.random.default_rng(seed)
    dirs = rng.normal(size=(num_points, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    radius = OBJECT_RADIUS * (_surface_radius(dirs, rng) + rng.normal(0.0, 0.01, size=num_points))
    xyz = dirs * radius[:, None]
    # screen coverage independent of P: scale ~ (1M / P)^(1/2)
    size_scale = OBJECT_RADIUS * (1.0e6 / num_points) ** 0.5
    log_s = rng.uniform(np.log(0.002), np.log(0.02), size=(num_points, 3))
    scaling = np.exp(log_s) * size_scale
    quat = rng.normal(size=(num_points, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    opacity = 1.0 / (1.0 + np.exp(-rng.normal(1.0, 2.0, size=(num_points, 1))))
    ncoef = (sh_degree + 1) ** 2
    feats = np.empty((num_points, 16, 3))
    feats[:, 0, :] = rng.uniform(-1.0, 1.5, size=(num_points, 3))
    feats[:, 1:, :] = rng.normal(0.0, 0.1, size=(num_points, 15, 3))
    feats[:, ncoef:, :] = 0.0
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return GaussianCloud(f32(xyz), f32(feats), f32(opacity), f32(scaling), f32(quat), sh_degree)


def look_at_euler_deg(position, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """Euler angles (deg) of the camera-to-world rotation, in the axis convention the
    reference's camera dicts use (x right, y up, z backward; `RT_from_rot_pos` flips y,z to
    get the OpenCV extrinsic, transformation_utils.py:34-35)."""
    position = np.asarray(position, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - position
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    rot_gl = np.stack([right, -down, -fwd], axis=1)  # columns: x right, y up, z backward
    return cam.matrix_to_euler_deg(rot_gl)


def ring_camera_positions(num_views: int, radius: float = CAMERA_RADIUS, elevation_deg: float = 20.0) -> np.ndarray:
    """"360 degree" capture: two rings at +/- elevation, num_views/2 each."""
    pos = []
    per_ring = max(1, (num_views + 1) // 2)
    for i in range(num_views):
        ring, k = divmod(i, per_ring)
        az = 2.0 * np.pi * (k + 0.5 * ring) / per_ring
        el = np.radians(elevation_deg if ring == 0 else -elevation_deg)
        pos.append([radius * np.cos(el) * np.cos(az), radius * np.cos(el) * np.sin(az), radius * np.sin(el)])
    return np.asarray(pos, dtype=np.float64)


def cap_camera_positions(num_views: int, radius: float = CAMERA_RADIUS, az_span_deg=120.0, el_span_deg=60.0) -> np.ndarray:
    """DTU-like frontal cap (config C2): a grid of cameras on az x el span."""
    cols = int(np.ceil(np.sqrt(num_views * az_span_deg / el_span_deg)))
    rows = int(np.ceil(num_views / cols))
    pos = []
    for i in range(num_views):
        r, c = divmod(i, cols)
        az = np.radians(-az_span_deg / 2 + az_span_deg * (c + 0.5) / cols)
        el = np.radians(-el_span_deg / 2 + el_span_deg * (r + 0.5) / rows)
        pos.append([radius * np.cos(el) * np.cos(az), radius * np.cos(el) * np.sin(az), radius * np.sin(el)])
    return np.asarray(pos, dtype=np.float64)


def make_stereo_cameras(num_views: int, width: int, height: int, layout: str = "ring", baseline_percentage: float = 7.0,
                        focal_factor: float = 0.9, principal_point=None):
    """Returns (cameras, baseline): `cameras` is the list of {'left','right'} dicts the
    reference Renderer builds (renderer_utils.py:178-206)."""
    if layout == "ring":
        positions = ring_camera_positions(num_views)
        scene_360 = True
    elif layout == "cap":
        positions = cap_camera_positions(num_views)
        scene_360 = False
    else:
        raise ValueError(f"unknown camera layout {layout!r}")
    baseline = cam.scene_baseline(positions, baseline_percentage, scene_360=scene_360)
    fx = fy = focal_factor * width
    cx, cy = (width / 2.0, height / 2.0) if principal_point is None else principal_point
    rigs = []
    for p in positions:
        rot = look_at_euler_deg(p)
        rigs.append(cam.make_stereo_rig(rot, tuple(p.tolist()), baseline, width, height, fx, fy, cx, cy))
    return rigs, baseline


# BASELINE.json `configs`, made concrete.  `tsdf_res` is the dense lattice
# resolution of the [-1,1]^3 window (voxel_length = 2 / tsdf_res; reference default 512).
CONFIGS = {
    "C0": dict(num_points=10_000, pairs=4, width=640, height=480, tsdf_res=128, layout="ring", min_db=4, max_db=20),
    "C1": dict(num_points=1_000_000, pairs=200, width=1600, height=1200, tsdf_res=512, layout="ring", min_db=4, max_db=20),
    "C2": dict(num_points=500_000, pairs=49, width=1600, height=1200, tsdf_res=512, layout="cap", min_db=4, max_db=20,
               principal_point=(823.2, 619.1)),
    "C3": dict(num_points=2_000_000, pairs=300, width=960, height=540, tsdf_res=768, layout="ring", min_db=4, max_db=15),
    "C4": dict(num_points=3_000_000, pairs=500, width=1920, height=1080, tsdf_res=1024, layout="ring", min_db=2, max_db=10),
}
