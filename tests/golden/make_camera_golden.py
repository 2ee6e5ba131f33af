"""Generates tests/golden/camera_golden.npz by calling the REFERENCE's own pose functions
(imported from /root/reference; only possible in the dev container, the vectors are committed).

  gs2mesh_utils/transformation_utils.py: eul2rotm, rotm2eul, RT_from_rot_pos, convert_R_T_to_GS,
                                          calculate_right_camera_pose
  third_party/gaussian-splatting/utils/graphics_utils.py: getWorld2View2, getProjectionMatrix
Run:  python tests/golden/make_camera_golden.py
"""
import os
import sys

import numpy as np

REF = os.environ.get("GS2MESH_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "third_party", "gaussian-splatting"))

from gs2mesh_utils.transformation_utils import (RT_from_rot_pos, calculate_right_camera_pose, convert_R_T_to_GS, eul2rotm,  # noqa: E402
                                                rotm2eul)
from utils.graphics_utils import getProjectionMatrix, getWorld2View2  # noqa: E402

rng = np.random.default_rng(1234)
N = 24
eulers = rng.uniform(-180, 180, size=(N, 3))
eulers[0] = [0, 0, 0]
eulers[1] = [0, 90, 0]  # gimbal branch of rotm2eul
eulers[2] = [10, -90, 30]
positions = rng.uniform(-3, 3, size=(N, 3))
baselines = rng.uniform(0.05, 0.3, size=N)
fovs = rng.uniform(0.3, 1.6, size=(N, 2))

out = dict(eulers=eulers, positions=positions, baselines=baselines, fovs=fovs)
out["rotm"] = np.stack([eul2rotm(e) for e in eulers])
out["eul_back"] = np.stack([rotm2eul(eul2rotm(e)) for e in eulers])
out["extrinsic"] = np.stack([RT_from_rot_pos(tuple(e), tuple(p)) for e, p in zip(eulers, positions)])
gs = [convert_R_T_to_GS(tuple(e), tuple(p)) for e, p in zip(eulers, positions)]
out["gs_R"] = np.stack([g[0] for g in gs])
out["gs_T"] = np.stack([g[1] for g in gs])
right = [calculate_right_camera_pose(np.asarray(e), tuple(p), b) for e, p, b in zip(eulers, positions, baselines)]
out["right_rot"] = np.stack([np.asarray(r[0]) for r in right])
out["right_pos"] = np.stack([np.asarray(r[1]) for r in right])
out["w2v"] = np.stack([getWorld2View2(g[0], g[1]) for g in gs])
out["proj"] = np.stack([getProjectionMatrix(0.01, 100.0, f[0], f[1]).numpy() for f in fovs])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "camera_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path)
