"""Known-answer tests of the marching-cubes restatement (oracle.extract_mesh_from_bricks)."""
import numpy as np

VL = 1.0 / 16
B0, NB = (-2, -2, -2), (4, 4, 4)


def _sphere_bricks(radius=0.55, trunc=0.2, hole=False):
    n = 64
    c = (np.arange(n) + 0.5) * VL + B0[0] * 16 * VL
    x, y, z = np.meshgrid(c, c, c, indexing="ij")
    sdf = np.sqrt(x * x + y * y + z * z) - radius  # negative inside, like a TSDF behind the surface seen from outside
    tsdf = np.clip(sdf / trunc, -1, 1).astype(np.float32)
    w = (np.abs(sdf) < trunc * 1.5).astype(np.float32)
    if hole:
        w[x > 0.3] = 0
    g = np.stack([tsdf * (w > 0), w], -1)
    tw = g.reshape(4, 16, 4, 16, 4, 16, 2).transpose(0, 2, 4, 1, 3, 5, 6).reshape(64, 4096, 2)
    return tw


def test_sphere_is_closed_and_on_the_isosurface(oracle):
    m = oracle.extract_mesh_from_bricks(_sphere_bricks(), B0, NB, VL)
    v, t = m["vertices"], m["triangles"]
    assert len(t) > 1000
    r = np.linalg.norm(v, axis=1)
    assert np.abs(r - 0.55).max() < 0.01  # linear interpolation error of a curved SDF
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    fw = {(a, b) for a, b in e}
    assert len(fw) == len(e) and all((b, a) in fw for a, b in e)  # closed, consistently oriented
    # orientation: Open3D's winding gives outward normals for tsdf < 0 inside
    n = np.cross(v[t[:, 1]] - v[t[:, 0]], v[t[:, 2]] - v[t[:, 0]])
    cen = v[t].mean(axis=1)
    assert (np.einsum("ij,ij->i", n, cen) > 0).mean() > 0.99
    # Euler characteristic of a sphere
    assert len(v) - len(fw) // 2 + len(t) == 2


def test_unobserved_voxels_cut_the_surface_open(oracle):
    m = oracle.extract_mesh_from_bricks(_sphere_bricks(hole=True), B0, NB, VL)
    assert m["vertices"][:, 0].max() < 0.3 + VL  # cubes with any weight-0 corner are skipped
    assert len(m["triangles"]) > 500


def test_vertex_lies_on_its_edge_with_open3d_interpolation(oracle):
    tw = _sphere_bricks()
    m = oracle.extract_mesh_from_bricks(tw, B0, NB, VL)
    g = tw.reshape(4, 4, 4, 16, 16, 16, 2).transpose(0, 3, 1, 4, 2, 5, 6).reshape(64, 64, 64, 2)[..., 0]
    for k, v in list(zip(m["keys"], m["vertices"]))[::97]:
        o = k[:3]
        d = np.zeros(3, int)
        d[k[3]] = 1
        f0, f1 = abs(g[tuple(o)]), abs(g[tuple(o + d)])
        want = (o + np.array(B0) * 16 + 0.5) * VL
        want[k[3]] += f0 * VL / (f0 + f1)
        np.testing.assert_allclose(v, want, atol=1e-12)
