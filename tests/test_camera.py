"""Host pose maths vs golden vectors produced by the reference's own functions
(tests/golden/make_camera_golden.py)."""
import os

import numpy as np
import pytest

from gs2mesh_b200 import camera as cam

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_golden.npz"))
N = len(G["eulers"])


def test_euler_matrix_roundtrip_matches_reference():
    for i in range(N):
        np.testing.assert_array_equal(cam.euler_deg_to_matrix(G["eulers"][i]), G["rotm"][i])
        np.testing.assert_array_equal(cam.matrix_to_euler_deg(cam.euler_deg_to_matrix(G["eulers"][i])), G["eul_back"][i])


def test_extrinsic_and_3dgs_pose_match_reference():
    for i in range(N):
        e, p = tuple(G["eulers"][i]), tuple(G["positions"][i])
        np.testing.assert_array_equal(cam.pose_c2w_opencv(e, p), G["extrinsic"][i])
        r, t = cam.pose_to_3dgs(e, p)
        np.testing.assert_array_equal(r, G["gs_R"][i])
        np.testing.assert_array_equal(t, G["gs_T"][i])


def test_right_camera_matches_reference():
    for i in range(N):
        rot, pos = cam.right_camera_pose(G["eulers"][i], tuple(G["positions"][i]), G["baselines"][i])
        np.testing.assert_array_equal(np.asarray(rot), G["right_rot"][i])
        np.testing.assert_array_equal(np.asarray(pos), G["right_pos"][i])


def test_view_and_projection_match_reference():
    for i in range(N):
        np.testing.assert_array_equal(cam.world_to_view(G["gs_R"][i], G["gs_T"][i]), G["w2v"][i])
        np.testing.assert_array_equal(cam.projection_matrix(0.01, 100.0, *G["fovs"][i]), G["proj"][i])


def test_render_and_tsdf_conventions_agree():
    """The view matrix used for rendering and inv('extrinsic') used by the TSDF stage describe the
    same camera (OpenCV axes), so rendered z is the TSDF depth."""
    rig = cam.make_stereo_rig((20.0, -35.0, 110.0), (1.0, -2.0, 0.5), 0.17, 640, 480, 576.0, 576.0, 320.0, 240.0)
    vt = cam.view_transforms_from_camera(rig["left"])
    w2c = np.linalg.inv(rig["left"]["extrinsic"])
    pts = np.random.default_rng(0).normal(size=(16, 3))
    a = (np.c_[pts, np.ones(16)] @ vt.world_view.astype(np.float64))[:, :3]
    b = (w2c @ np.c_[pts, np.ones(16)].T).T[:, :3]
    np.testing.assert_allclose(a, b, atol=2e-6)
    # right camera sits `baseline` along the left camera's +x
    right_c = np.asarray(rig["right"]["pos"])
    np.testing.assert_allclose((w2c @ np.r_[right_c, 1.0])[:3], [0.17, 0, 0], atol=1e-6)
    vt_r = cam.view_transforms_from_camera(rig["right"])
    np.testing.assert_allclose(vt_r.world_view[:3, :3], vt.world_view[:3, :3], atol=1e-6)


def test_packed_record_layout():
    rig = cam.make_stereo_rig((0, 0, 0), (0, 0, 3), 0.1, 64, 48, 50.0, 50.0, 32.0, 24.0)
    vt = cam.view_transforms_from_camera(rig["left"])
    rec = vt.packed()
    assert rec.shape == (cam.CAMERA_RECORD_FLOATS,) and rec.dtype == np.float32
    np.testing.assert_array_equal(rec[:16].reshape(4, 4), vt.world_view)
    np.testing.assert_array_equal(rec[16:32].reshape(4, 4), vt.full_proj)
    np.testing.assert_array_equal(rec[32:35], vt.cam_center)
    assert vt.tan_fovx == pytest.approx(64 / (2 * 50.0)) and vt.tan_fovy == pytest.approx(48 / (2 * 50.0))


def test_baseline_rule():
    from gs2mesh_b200 import scene

    pos = scene.ring_camera_positions(40)
    b = cam.scene_baseline(pos, 7.0, scene_360=True)
    assert b == pytest.approx(0.07 * scene.CAMERA_RADIUS, rel=1e-6)
    assert cam.scene_baseline(pos, 7.0, scene_360=True, dtu_compat=True) == pytest.approx(2 * b)
    cap = scene.cap_camera_positions(49)
    assert cam.scene_baseline(cap, 7.0, scene_360=False) == pytest.approx(0.07 * scene.CAMERA_RADIUS, rel=1e-4)
