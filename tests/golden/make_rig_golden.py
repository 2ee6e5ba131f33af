"""Generates tests/golden/rig_golden.json + tests/golden/colmap_fixture/*: stereo camera rigs produced by the
REFERENCE's own `Renderer.__init__` (gs2mesh_utils/renderer_utils.py:106-216) on small synthetic COLMAP text
models.  The reference module imports plotting / Open3D / trimesh packages that are absent here; they are
stubbed (none of them is touched by the constructor's pose code).  Dev-container only; outputs are committed.
Run:  python tests/golden/make_rig_golden.py
"""
import faulthandler
import json
faulthandler.dump_traceback_later(90, exit=True)
import os
import sys
import types
from argparse import Namespace

import numpy as np
from scipy.spatial.transform import Rotation

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GS2MESH_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gs2mesh_b200", "compat"))
sys.path.insert(0, REF)


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


for pkg in ("matplotlib", "plotly"):
    stub(pkg).__path__ = []
for name in ("matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "plotly.graph_objects", "plotly.graph_objs", "plotly.io",
             "plotly.subplots", "plotly.express", "open3d", "simple_knn", "k3d", "trimesh", "mediapy", "imageio"):
    stub(name)
stub("plyfile", PlyData=object, PlyElement=object)
stub("simple_knn._C", distCUDA2=None)

import gs2mesh_utils.renderer_utils as ru  # noqa: E402

ru.read_ply = lambda path: (None, None)  # Open3D point-cloud read, only used for visualisation

from gs2mesh_b200 import scene  # noqa: E402

CASES = {
    "ring360": dict(n=10, layout="ring", model="PINHOLE", shared_camera=True, scene_360=True, sort=False, dataset="custom", pct=7.0),
    "cap": dict(n=9, layout="cap", model="SIMPLE_RADIAL", shared_camera=False, scene_360=False, sort=False, dataset="custom", pct=14.0),
    # renderer_sort_cameras is not exercised: the reference's sort (renderer_utils.py:34-99) does not terminate when its
    # 2-nearest candidate list runs into an already visited camera, which happens on these fixtures
    "dtu": dict(n=6, layout="ring", model="PINHOLE", shared_camera=True, scene_360=True, sort=False, dataset="DTU", pct=7.0),
}


def write_fixture(dirname, cfg, seed):
    rng = np.random.default_rng(seed)
    W, H = 640, 480
    rigs, _ = scene.make_stereo_cameras(cfg["n"], W, H, layout=cfg["layout"])
    os.makedirs(os.path.join(dirname, "sparse", "0"), exist_ok=True)
    order = rng.permutation(cfg["n"])  # image ids not in spatial order
    with open(os.path.join(dirname, "sparse", "0", "images.txt"), "w") as f:
        f.write("# Image list with two lines of data per image:\n")
        for image_id, k in enumerate(order, start=1):
            c2w = rigs[k]["left"]["extrinsic"].copy()
            c2w[:3, 3] += rng.normal(0, 0.02, 3)
            w2c = np.linalg.inv(c2w)
            q = Rotation.from_matrix(w2c[:3, :3]).as_quat()  # x y z w
            qvec = [q[3], q[0], q[1], q[2]]
            cam_id = 1 if cfg["shared_camera"] else image_id
            f.write(" ".join([str(image_id)] + [repr(float(v)) for v in qvec] + [repr(float(v)) for v in w2c[:3, 3]] +
                             [str(cam_id), f"img_{image_id:03d}.png"]) + "\n")
            f.write("10.5 20.25 -1 30.0 40.0 7\n")
    with open(os.path.join(dirname, "sparse", "0", "cameras.txt"), "w") as f:
        f.write("# Camera list with one line of data per camera:\n")
        ncam = 1 if cfg["shared_camera"] else cfg["n"]
        for cam_id in range(1, ncam + 1):
            fx = 0.9 * W + cam_id
            if cfg["model"] == "PINHOLE":
                f.write(f"{cam_id} PINHOLE {W} {H} {fx} {fx + 2.5} {W / 2 + 0.25} {H / 2 - 0.75}\n")
            else:
                f.write(f"{cam_id} SIMPLE_RADIAL {W} {H} {fx} {W / 2 + 0.25} {H / 2 - 0.75} 0.01\n")


def jsonable(o):
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    return o


golden = {}
for seed, (name, cfg) in enumerate(CASES.items()):
    d = os.path.join(HERE, "colmap_fixture", name)
    write_fixture(d, cfg, seed)
    args = Namespace(colmap_name=name, GS_white_background=False, GS_iterations=30000, renderer_baseline_absolute=None,
                     renderer_scene_360=cfg["scene_360"], dataset_name=cfg["dataset"], renderer_baseline_percentage=cfg["pct"],
                     renderer_sort_cameras=cfg["sort"], renderer_save_json=False)
    r = ru.Renderer("/nonexistent_base", d, "/nonexistent_out", args, dataset=cfg["dataset"])
    golden[name] = dict(args={k: v for k, v in vars(args).items()}, baseline=float(r.baseline),
                        sorted_camera_indices=[int(v) for v in r.sorted_camera_indices], cameras=jsonable(r.cameras),
                        poses=jsonable(np.asarray(r.poses)))
with open(os.path.join(HERE, "rig_golden.json"), "w") as f:
    json.dump(golden, f)
print("wrote rig_golden.json", {k: len(v["cameras"]) for k, v in golden.items()})
