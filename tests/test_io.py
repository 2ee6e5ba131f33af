"""Loaders (COLMAP text model -> stereo rigs, 3DGS point_cloud.ply -> activated tensors) against golden rigs
produced by the reference's own Renderer.__init__ (tests/golden/make_rig_golden.py)."""
import json
import os
from argparse import Namespace

import numpy as np
import pytest

from gs2mesh_b200 import io as gio
from gs2mesh_b200 import scene

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "rig_golden.json")))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_stereo_rigs_match_reference_constructor(name):
    g = GOLD[name]
    d = os.path.join(HERE, "golden", "colmap_fixture", name)
    args = Namespace(**g["args"])
    poses = gio.poses_from_file(os.path.join(d, "sparse", "0", "images.txt"))
    np.testing.assert_array_equal(poses, np.asarray(g["poses"]))
    rigs, baseline, order = gio.build_stereo_rigs(poses, gio.read_cameras_text(os.path.join(d, "sparse", "0", "cameras.txt")), args)
    assert baseline == g["baseline"]
    assert list(order) == g["sorted_camera_indices"]
    assert len(rigs) == len(g["cameras"])
    for mine, ref in zip(rigs, g["cameras"]):
        for side in ("left", "right"):
            for k in ("rot", "pos"):
                assert list(mine[side][k]) == list(ref[side][k]), (side, k)
            for k in ("width", "height", "fx", "fy", "cx", "cy"):
                assert mine[side][k] == ref[side][k], (side, k)
            np.testing.assert_array_equal(np.asarray(mine[side]["intrinsic"]), np.asarray(ref[side]["intrinsic"]))
            np.testing.assert_array_equal(np.asarray(mine[side]["extrinsic"]), np.asarray(ref[side]["extrinsic"]))
        assert mine["left"]["baseline"] == ref["left"]["baseline"]


def test_camera_sort_terminates_and_chains_neighbours():
    # a helix: the natural order is by height
    t = np.linspace(0, 4 * np.pi, 12)
    pts = np.stack([np.cos(t), np.sin(t), 0.2 * t], 1)
    rng = np.random.default_rng(0)
    perm = rng.permutation(12)
    order = gio.sort_camera_coordinates(pts[perm])
    assert sorted(order) == list(range(12))
    assert perm[order[0]] == 0  # starts at the lowest camera
    # two rings (the fixture on which the reference's loop does not terminate) still yields a permutation
    ring = scene.ring_camera_positions(8)
    assert sorted(gio.sort_camera_coordinates(ring)) == list(range(8))


def test_gaussian_ply_roundtrip(tmp_path):
    cloud = scene.make_gaussians(500, seed=4)
    path = str(tmp_path / "point_cloud.ply")
    gio.write_gaussian_ply(path, cloud)
    back = gio.read_gaussian_ply(path)
    np.testing.assert_array_equal(back.xyz, cloud.xyz)
    np.testing.assert_array_equal(back.features, cloud.features)  # incl. the channel-major f_rest layout
    np.testing.assert_allclose(back.opacity, cloud.opacity, rtol=2e-6)
    np.testing.assert_allclose(back.scaling, cloud.scaling, rtol=2e-6)
    np.testing.assert_allclose(back.rotation, cloud.rotation, atol=1e-6)
    assert back.opacity.shape == (500, 1) and back.features.shape == (500, 16, 3)
    with pytest.raises(ValueError, match="f_rest"):
        gio.read_gaussian_ply(path, sh_degree=2)


def test_load_reference_scene_layout(tmp_path):
    """Directory layout of renderer_utils.py:118-129: <base>/splatting_output/<splatting>/<colmap_name>/point_cloud/
    iteration_<N>/point_cloud.ply and <colmap_dir>/sparse/0/{images,cameras}.txt."""
    name = "ring360"
    g = GOLD[name]
    args = Namespace(**g["args"])
    cloud = scene.make_gaussians(64, seed=1)
    gio.write_gaussian_ply(str(tmp_path / "splatting_output" / "custom" / name / "point_cloud" / "iteration_30000" / "point_cloud.ply"), cloud)
    rigs, baseline, loaded, poses, order = gio.load_reference_scene(str(tmp_path), os.path.join(HERE, "golden", "colmap_fixture", name),
                                                                   args, "custom")
    assert len(rigs) == 10 and baseline == g["baseline"] and loaded.num_points == 64 and poses.shape == (10, 3, 4)
    from gs2mesh_b200.renderer import Renderer

    r = Renderer(str(tmp_path), os.path.join(HERE, "golden", "colmap_fixture", name), None, args, dataset="custom", splatting="custom")
    assert len(r) == 10 and r.baseline == g["baseline"] and r.left_cameras[3]["fx"] == g["cameras"][3]["left"]["fx"]
