"""The Open3D-shaped facade (gs2mesh_b200/o3d_compat.py): CPU tests of the host-only members, GPU test replaying
gs2mesh_utils/tsdf_utils.py:53-142 call for call against the oracle volume."""
import os

import numpy as np
import pytest


def test_module_tree_matches_the_calls_tsdf_utils_makes():
    import gs2mesh_b200.o3d_compat as o3d

    # every attribute chain tsdf_utils.py dereferences (:53-56, :88-93, :106-110, :119, :132-140)
    assert o3d.pipelines.integration.ScalableTSDFVolume
    assert o3d.pipelines.integration.TSDFVolumeColorType.RGB8 == 1
    assert o3d.geometry.Image and o3d.geometry.RGBDImage.create_from_color_and_depth
    assert o3d.camera.PinholeCameraIntrinsic
    assert o3d.io.write_triangle_mesh
    with o3d.utility.VerbosityContextManager(o3d.utility.VerbosityLevel.Debug):
        pass


def test_rgbd_and_intrinsic_host_objects():
    import gs2mesh_b200.o3d_compat as o3d

    rgb = np.zeros((4, 6, 3), np.uint8)
    depth = np.ones((4, 6), np.float32)
    rgbd = o3d.geometry.RGBDImage.create_from_color_and_depth(o3d.geometry.Image(rgb), o3d.geometry.Image(depth), depth_scale=2.0,
                                                               depth_trunc=5.0, convert_rgb_to_intensity=False)
    assert np.asarray(rgbd.depth).shape == (4, 6) and np.asarray(rgbd.color).dtype == np.uint8
    assert rgbd.depth.width == 6 and rgbd.depth.height == 4
    with pytest.raises(RuntimeError, match="Unsupported image format"):
        o3d.geometry.RGBDImage.create_from_color_and_depth(o3d.geometry.Image(rgb), o3d.geometry.Image(depth[:3]),
                                                           convert_rgb_to_intensity=False)
    with pytest.raises(NotImplementedError):
        o3d.geometry.RGBDImage.create_from_color_and_depth(o3d.geometry.Image(rgb), o3d.geometry.Image(depth))
    k = o3d.camera.PinholeCameraIntrinsic(6, 4, 10.0, 11.0, 3.0, 2.0)
    np.testing.assert_array_equal(k.intrinsic_matrix, [[10, 0, 3], [0, 11, 2], [0, 0, 1]])
    assert k.get_focal_length() == (10.0, 11.0) and k.get_principal_point() == (3.0, 2.0)


def test_mesh_mask_removal_and_ply(tmp_path):
    import gs2mesh_b200.o3d_compat as o3d
    from gs2mesh_b200.io import read_gaussian_ply  # noqa: F401  (package import must not need a GPU)

    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5], [6, 5, 5], [5, 6, 5], [9, 9, 9]], float)
    t = np.array([[0, 1, 2], [3, 4, 5]])
    m = o3d.geometry.TriangleMesh(v, t, vertex_colors=np.linspace(0, 1, 21).reshape(7, 3))
    labels, counts, area = m.cluster_connected_triangles()
    assert sorted(counts.tolist()) == [1, 1] and np.allclose(area, 0.5)
    m.remove_triangles_by_mask(np.array([False, True]))
    m.remove_unreferenced_vertices()
    assert m.vertices.shape == (3, 3) and m.triangles.tolist() == [[0, 1, 2]] and m.vertex_colors.shape == (3, 3)
    path = os.path.join(tmp_path, "m.ply")
    assert o3d.io.write_triangle_mesh(path, m)
    head = open(path, "rb").read(600)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0") and b"element vertex 3" in head and b"element face 1" in head


@pytest.mark.gpu
def test_tsdf_utils_call_sequence_against_oracle(oracle, gsb_lib, cuda_device, tmp_path):
    """tsdf_utils.py:53-142 with `o3d` bound to the facade, vs the oracle's Open3D restatement."""
    import copy

    import torch

    import gs2mesh_b200.o3d_compat as o3d
    from tests.test_gpu_tsdf import CX, CY, FX, FY, H, TRUNC, VL, W, _views

    scale, trunc = 1.0, 4.0
    with torch.cuda.device(cuda_device):
        volume = o3d.pipelines.integration.ScalableTSDFVolume(voxel_length=VL, sdf_trunc=TRUNC,
                                                              color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8,
                                                              window_resolution=128, device=cuda_device)  # :53-56
        ovol = oracle.OracleTSDFVolume(VL, TRUNC, with_color=True)
        for depth, rgb, w2c in _views(4):
            depth = depth.copy()
            depth[depth < 0.5] = 0  # :83
            rgbd = o3d.geometry.RGBDImage.create_from_color_and_depth(o3d.geometry.Image(rgb), o3d.geometry.Image(depth),
                                                                       depth_scale=scale, depth_trunc=trunc,
                                                                       convert_rgb_to_intensity=False)  # :88-93
            intrinsic = o3d.camera.PinholeCameraIntrinsic(W, H, FX, FY, CX, CY)  # :106
            c2w = np.linalg.inv(w2c)
            volume.integrate(rgbd, intrinsic, np.linalg.inv(c2w))  # :107
            ovol.integrate(depth, rgb, W, H, FX, FY, CX, CY, w2c, depth_scale=scale, depth_trunc=trunc)
        torch.cuda.synchronize()
        tw, alloc, n_out = ovol.export_bricks(volume._vol.brick_origin, volume._vol.brick_count)
        assert n_out == 0
        np.testing.assert_array_equal(volume._vol.bricks().cpu().numpy(), tw)

        mesh = volume.extract_triangle_mesh()  # :108
        mesh.scale(0.5, (0, 0, 0))  # :109
        mesh.compute_vertex_normals()  # :110
        ref = oracle.extract_mesh_from_bricks(tw, volume._vol.brick_origin, volume._vol.brick_count, VL)
        assert len(mesh.triangles) == len(ref["triangles"]) > 1000
        assert mesh.vertex_normals.shape == mesh.vertices.shape
        np.testing.assert_allclose(np.sort(mesh.vertices[:, 0]), np.sort(ref["vertices"][:, 0] * 0.5), rtol=0, atol=1e-6)
        o3d.io.write_triangle_mesh(os.path.join(tmp_path, "x_mesh.ply"), mesh)  # :119

        mesh_0 = copy.deepcopy(mesh)  # :131
        with o3d.utility.VerbosityContextManager(o3d.utility.VerbosityLevel.Debug):
            tri_clusters, n_tri, _area = mesh_0.cluster_connected_triangles()  # :132-133
        tri_clusters, n_tri = np.asarray(tri_clusters), np.asarray(n_tri)
        mask = n_tri[tri_clusters] < 100  # :137
        mesh_0.remove_triangles_by_mask(mask)  # :138
        mesh_0.remove_unreferenced_vertices()
        assert 0 < len(mesh_0.triangles) <= len(mesh.triangles)
        assert mesh_0.triangles.max() == len(mesh_0.vertices) - 1
        assert len(mesh.triangles) == len(ref["triangles"])  # the deep copy left the original alone

    with pytest.raises(RuntimeError, match="Unsupported image format"):
        bad = o3d.geometry.RGBDImage.create_from_color_and_depth(o3d.geometry.Image(rgb[:, :10]), o3d.geometry.Image(depth[:, :10]),
                                                                 convert_rgb_to_intensity=False)
        volume.integrate(bad, intrinsic, np.eye(4))


def test_stage_view_selection_accepts_lists_and_strings():
    """tsdf_utils.py:58-63: TSDF_dilate / TSDF_valid / TSDF_skip pick the views; the reference declares the two lists as
    strings (argument_utils.py:76-77), so both spellings are accepted."""
    from types import SimpleNamespace

    from gs2mesh_b200.tsdf import TSDF

    rend = SimpleNamespace(device="cuda", baseline=0.1)
    mk = lambda **kw: TSDF(rend, None, SimpleNamespace(**kw), "t")
    assert [i for i in range(8) if mk(TSDF_dilate=3)._selected(i)] == [0, 3, 6]
    assert [i for i in range(8) if mk(TSDF_valid=[1, 2, 5])._selected(i)] == [1, 2, 5]
    assert [i for i in range(8) if mk(TSDF_valid="[1, 2, 5]", TSDF_skip="2")._selected(i)] == [1, 5]
    assert [i for i in range(6) if mk(TSDF_skip=(0, 4), TSDF_dilate=2)._selected(i)] == [2]
    assert all(mk()._selected(i) for i in range(4))
