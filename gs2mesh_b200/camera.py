"""Host-side camera maths for the stereo renderer (numpy, no GPU).

Mirrors the reference's pose conventions so that a camera dict produced by the
reference `Renderer.__init__` renders identically here:

* Euler <-> matrix, left->right stereo pose, COLMAP->3DGS pose conversion:
  gs2mesh_utils/transformation_utils.py:23-63, 79-135, 207-224
* world->view / projection matrices and the packed per-view transforms the
  rasterizer consumes: third_party/gaussian-splatting/utils/graphics_utils.py:38-71,
  third_party/gaussian-splatting/scene/cameras.py:51-57
* FoV from intrinsics (principal point deliberately ignored, as the reference does):
  gs2mesh_utils/renderer_utils.py:384-385

Everything is checked against golden vectors generated from the reference's own
functions (tests/golden/make_camera_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

_TINY = 1e-7  # transformation_utils.py:15-16 (`fix_zero`)

ZNEAR = 0.01  # scene/cameras.py:48-49
ZFAR = 100.0


def snap_zero(a):
    """|a| < 1e-7 -> 0 (transformation_utils.py:16)."""
    a = np.asarray(a)
    return np.where(np.abs(a) < _TINY, 0, a)


def euler_deg_to_matrix(angles_deg) -> np.ndarray:
    """R = Rz @ Ry @ Rx built in float32 from XYZ Euler angles in degrees
    (transformation_utils.py:79-109)."""
    ax, ay, az = np.radians(angles_deg)
    cx, sx = np.cos(ax), np.sin(ax)
    cy, sy = np.cos(ay), np.sin(ay)
    cz, sz = np.cos(az), np.sin(az)
    rot_x = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float32)
    rot_y = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float32)
    rot_z = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float32)
    return snap_zero(rot_z @ rot_y @ rot_x)


def matrix_to_euler_deg(rot) -> np.ndarray:
    """Inverse of `euler_deg_to_matrix` incl. the gimbal branch
    (transformation_utils.py:111-135)."""
    m = np.asarray(rot, dtype=np.float32)
    sy = np.sqrt(m[0, 0] ** 2 + m[1, 0] ** 2)
    if sy < 1e-6:
        ex = np.arctan2(-m[1, 2], m[1, 1])
        ey = np.arctan2(-m[2, 0], sy)
        ez = 0
    else:
        ex = np.arctan2(m[2, 1], m[2, 2])
        ey = np.arctan2(-m[2, 0], sy)
        ez = np.arctan2(m[1, 0], m[0, 0])
    return snap_zero(np.degrees([ex, ey, ez]))


def pose_c2w_opencv(rot_deg, pos) -> np.ndarray:
    """4x4 camera-to-world with the y/z axes flipped to OpenCV convention -- the
    `'extrinsic'` entry of a camera dict (transformation_utils.py:23-40)."""
    pose = np.eye(4)
    rot = euler_deg_to_matrix(rot_deg)
    rot[:, 1:] *= -1
    pose[:3, :3] = rot
    pose[:3, 3] = np.array(pos)
    return pose


def pose_to_3dgs(rot_deg, pos):
    """(R, T) in the convention 3DGS `Camera` expects (transformation_utils.py:42-63)."""
    c2w = np.zeros((4, 4))
    c2w[:3, :3] = euler_deg_to_matrix(rot_deg)
    c2w[:3, 3] = np.asarray(pos, dtype=np.float32)
    c2w[3, 3] = 1.0
    w2c = np.linalg.inv(c2w)
    t = w2c[:3, 3]
    t[1:] *= -1
    r = w2c[:3, :3].transpose()
    r[:, 1:] *= -1
    return r, t


def right_camera_pose(rot_left_deg, pos_left, baseline):
    """Right camera = left camera shifted by `baseline` along its +x axis
    (transformation_utils.py:207-224)."""
    rot_left_deg = np.asarray(rot_left_deg)
    shift = euler_deg_to_matrix(rot_left_deg) @ np.array([baseline, 0, 0], dtype=np.float32)
    pos_right = np.array(pos_left, dtype=np.float32) + shift
    return tuple(rot_left_deg.tolist()), tuple(snap_zero(pos_right).tolist())


def intrinsic_matrix(cam) -> np.ndarray:
    """3x3 K from a dict with fx, fy, cx, cy (transformation_utils.py:65-77)."""
    return np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]])


def world_to_view(r, t, translate=(0.0, 0.0, 0.0), scale=1.0) -> np.ndarray:
    """float32 4x4 world->view (graphics_utils.py:38-49)."""
    rt = np.zeros((4, 4))
    rt[:3, :3] = np.asarray(r).transpose()
    rt[:3, 3] = t
    rt[3, 3] = 1.0
    c2w = np.linalg.inv(rt)
    c2w[:3, 3] = (c2w[:3, 3] + np.asarray(translate)) * scale
    return np.float32(np.linalg.inv(c2w))


def projection_matrix(znear, zfar, fov_x, fov_y) -> np.ndarray:
    """float32 4x4 perspective matrix, z mapped to [0,1], +z forward
    (graphics_utils.py:51-71)."""
    tan_y = math.tan(fov_y / 2)
    tan_x = math.tan(fov_x / 2)
    top, right = tan_y * znear, tan_x * znear
    bottom, left = -top, -right
    p = np.zeros((4, 4), dtype=np.float32)
    p[0, 0] = 2.0 * znear / (right - left)
    p[1, 1] = 2.0 * znear / (top - bottom)
    p[0, 2] = (right + left) / (right - left)
    p[1, 2] = (top + bottom) / (top - bottom)
    p[3, 2] = 1.0
    p[2, 2] = zfar / (zfar - znear)
    p[2, 3] = -(zfar * znear) / (zfar - znear)
    return p


def fov_from_focal(size_px, focal_px) -> float:
    """renderer_utils.py:384-385: 2*atan2(size, 2*f) (principal point ignored)."""
    return float(2 * np.arctan2(size_px, 2 * focal_px))


CAMERA_RECORD_FLOATS = 36  # view(16) | proj(16) | campos(3) | pad -> records stay 16-byte aligned


@dataclass
class ViewTransforms:
    """What `scene/cameras.py:54-57` leaves on the GPU for one view, as float32 host
    arrays.  `world_view` and `full_proj` are stored TRANSPOSED (row-vector
    convention), i.e. exactly the memory the rasterizer kernels index as
    matrix[0],[4],[8],[12] (auxiliary.h:58-77)."""

    world_view: np.ndarray  # [4,4] f32
    full_proj: np.ndarray  # [4,4] f32
    cam_center: np.ndarray  # [3] f32
    tan_fovx: float
    tan_fovy: float
    width: int
    height: int

    def packed(self) -> np.ndarray:
        """The 36-float device-side camera record: view(16) | proj(16) | campos(3) | 0."""
        rec = np.zeros(CAMERA_RECORD_FLOATS, dtype=np.float32)
        rec[0:16] = self.world_view.reshape(-1)
        rec[16:32] = self.full_proj.reshape(-1)
        rec[32:35] = self.cam_center.reshape(-1)
        return rec


def view_transforms(r, t, fov_x, fov_y, width, height) -> ViewTransforms:
    """Equivalent of constructing `scene.cameras.Camera` (cameras.py:51-57) without the
    dummy image upload (renderer_utils.py:386)."""
    wv = np.ascontiguousarray(world_to_view(r, t).transpose(), dtype=np.float32)
    proj = np.ascontiguousarray(projection_matrix(ZNEAR, ZFAR, fov_x, fov_y).transpose(), dtype=np.float32)
    full = (wv @ proj).astype(np.float32)
    center = np.linalg.inv(wv)[3, :3].astype(np.float32)
    return ViewTransforms(
        world_view=wv,
        full_proj=np.ascontiguousarray(full),
        cam_center=np.ascontiguousarray(center),
        tan_fovx=math.tan(fov_x * 0.5),  # gaussian_renderer/__init__.py:33-34
        tan_fovy=math.tan(fov_y * 0.5),
        width=int(width),
        height=int(height),
    )


def view_transforms_from_camera(cam: dict) -> ViewTransforms:
    """Camera dict (renderer_utils.py:182-206) -> per-view transforms: the body of the
    loop in renderer_utils.py:379-386."""
    r, t = pose_to_3dgs(tuple(cam["rot"]), tuple(cam["pos"]))
    w, h = cam["width"], cam["height"]
    return view_transforms(r, t, fov_from_focal(w, cam["fx"]), fov_from_focal(h, cam["fy"]), w, h)


def make_stereo_rig(rot_deg, pos, baseline, width, height, fx, fy, cx, cy) -> dict:
    """One {'left','right'} camera pair as built in renderer_utils.py:178-206 (incl. the
    quirk that the right camera's 'extrinsic' is the LEFT camera's, :204)."""
    # dtype matters: the reference keeps the float32 Euler angles `rotm2eul` returns for 'extrinsic' and the right
    # camera (float32 radians inside eul2rotm) but stores python floats in 'rot' (renderer_utils.py:180-204)
    rot_arr = np.asarray(rot_deg)
    pos = tuple(float(v) for v in pos)
    rot_r, pos_r = right_camera_pose(rot_arr, pos, baseline)
    k = {"fx": fx, "fy": fy, "cx": cx, "cy": cy}
    common = {"width": int(width), "height": int(height), "fx": float(fx), "fy": float(fy), "cx": float(cx), "cy": float(cy)}
    left = dict(rot=tuple(rot_arr.tolist()), pos=pos, **common, intrinsic=intrinsic_matrix(k),
                extrinsic=pose_c2w_opencv(tuple(rot_arr), pos), baseline=baseline)
    right = dict(rot=rot_r, pos=pos_r, **common, intrinsic=intrinsic_matrix(k),
                 extrinsic=pose_c2w_opencv(tuple(rot_arr), pos))
    return {"left": left, "right": right}


def scene_baseline(camera_locations, percentage=7.0, scene_360=True, dtu_compat=False) -> float:
    """Stereo baseline = percentage of the scene radius (renderer_utils.py:154-170)."""
    ts = np.array(camera_locations)
    if scene_360:
        radius = np.median(np.linalg.norm(ts - ts.mean(axis=0), axis=1))
        if dtu_compat:
            radius *= 2
    else:
        from scipy.optimize import least_squares

        x_m, y_m, z_m = np.mean(ts, axis=0)
        guess = np.array([x_m, y_m, z_m, 1.0])
        x, y, z = ts[:, 0], ts[:, 1], ts[:, 2]
        fit = least_squares(lambda p: np.sqrt((x - p[0]) ** 2 + (y - p[1]) ** 2 + (z - p[2]) ** 2) - p[3], guess)
        radius = fit.x[3]
    return float(radius * (percentage / 100))
