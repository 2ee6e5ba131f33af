#!/usr/bin/env python
"""Generates tests/golden/tsdf_o3d_<case>.npz from the REAL wheel the reference's TSDF half is: open3d==0.17.0
(/root/reference/requirements.txt:15).  It runs the literal call sequence of gs2mesh_utils/tsdf_utils.py

    :53-56   volume = o3d.pipelines.integration.ScalableTSDFVolume(voxel_length, sdf_trunc, color_type=RGB8)
    :85-86   extrinsic[:3, 3] /= TSDF_scale
    :88-93   rgbd = o3d.geometry.RGBDImage.create_from_color_and_depth(Image(rgb), Image(depth), depth_scale=TSDF_scale,
                                                                      depth_trunc=..., convert_rgb_to_intensity=False)
    :106-107 volume.integrate(rgbd, PinholeCameraIntrinsic(w, h, fx, fy, cx, cy), inv(extrinsic))
    :108-110 mesh = volume.extract_triangle_mesh(); mesh.scale(TSDF_scale, (0,0,0)); mesh.compute_vertex_normals()

on seeded synthetic frames (no GPU, no dataset) and dumps what the Python API lets one read back:
    extract_voxel_point_cloud()  -> voxel centres with weight != 0 and |tsdf| < 0.98, colour = (tsdf + 1) / 2  (fp64)
    extract_triangle_mesh()      -> vertices / triangles / vertex_colors (before and after scale), vertex normals
The inputs travel inside the .npz, so the consuming tests (tests/test_oracle_tsdf.py::test_open3d_golden_*,
tests/test_oracle_mesh.py::test_open3d_golden_mesh) need neither open3d nor this script.

STATUS (round 2): the wheel could not be obtained -- no network in the dev container or on the GPU box, no copy in
/opt/wheelhouse, and open3d 0.17.0 publishes no cp312 wheel (this image is Python 3.12.3).  Transcripts:
profiles/r02_open3d_attempt_devbox.log, profiles/r02_open3d_attempt_gpubox.log.  The fixtures therefore do not exist yet and
the consuming tests skip with that reason; the TSDF/mesh oracle stays "parity unpinned".

    python tests/golden/make_tsdf_golden.py                  # needs `import open3d` == 0.17.0
    python tests/golden/make_tsdf_golden.py --backend oracle --out /tmp/x   # plumbing check of the consuming tests ONLY:
                                                                            # writes the same file format from the CPU oracle;
                                                                            # never commit those files as golden
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> frame geometry + the reference flags that reach Open3D (argument_utils.py:74-90)
CASES = {
    # centred principal point, TSDF_scale = 1 (custom / DTU default)
    "centred": dict(W=160, H=120, fx=150.0, fy=150.0, cx=80.0, cy=60.0, scale=1.0, voxel=8.0, trunc=0.06, views=3, seed=11,
                    centre=(0.0, 0.0, 0.0)),
    # DTU-like off-centre principal point (run_single.py:126 passes cx, cy = 823.2, 619.1 at 1600x1200)
    "offcentre": dict(W=160, H=120, fx=150.0, fy=145.0, cx=82.32, cy=61.91, scale=1.0, voxel=8.0, trunc=0.06, views=3, seed=12,
                      centre=(0.0, 0.0, 0.0)),
    # TSDF_scale = 0.1 (MobileBrick default, argument_utils.py:35) and a scene far from the origin: the unbounded volume
    "scaled_shifted": dict(W=160, H=120, fx=150.0, fy=150.0, cx=80.0, cy=60.0, scale=0.1, voxel=32.0, trunc=0.25, views=4, seed=13,
                           centre=(5.0, -3.0, 2.0)),
}
MIN_DEPTH_BASELINES, MAX_DEPTH_BASELINES, BASELINE = 4.0, 20.0, 0.16


def look_at_c2w(pos, target):
    """Camera-to-world, OpenCV axes (x right, y down, z forward)."""
    fwd = np.asarray(target, float) - np.asarray(pos, float)
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, down, fwd, pos
    return m


def make_frames(case):
    """Seeded synthetic views of a bumpy sphere of radius 0.5 around `centre`: float32 depth (z along the optical axis,
    0 where the ray misses), uint8 rgb, 4x4 camera-to-world extrinsic -- what tsdf_utils.py:65-67,85 reads per view."""
    c = CASES[case]
    rng = np.random.default_rng(c["seed"])
    W, H = c["W"], c["H"]
    centre = np.asarray(c["centre"], float)
    frames = []
    for k in range(c["views"]):
        az = 2 * np.pi * (k + 0.3) / c["views"]
        pos = centre + 2.0 * np.array([np.cos(az) * 0.94, np.sin(az) * 0.94, 0.34])
        c2w = look_at_c2w(pos, centre)
        v, u = np.mgrid[0:H, 0:W]
        d_cam = np.stack([(u - c["cx"]) / c["fx"], (v - c["cy"]) / c["fy"], np.ones((H, W))], -1)  # z = 1 rays
        d_w = d_cam @ c2w[:3, :3].T
        o = pos - centre
        # |o + t d|^2 = r^2 with a direction-dependent radius (two Newton refinements from the plain sphere)
        r0 = 0.5
        a = (d_w ** 2).sum(-1)
        b = 2 * (d_w @ o)
        t = np.zeros((H, W))
        hit = np.zeros((H, W), bool)
        for _ in range(3):
            disc = b * b - 4 * a * (o @ o - r0 ** 2)
            hit = disc > 0
            t = np.where(hit, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 0)
            p = o + t[..., None] * d_w
            n = p / np.maximum(np.linalg.norm(p, axis=-1, keepdims=True), 1e-9)
            r0 = 0.5 + 0.04 * np.cos(3 * np.arccos(np.clip(n[..., 2], -1, 1))) + 0.03 * n[..., 0] * n[..., 1]
        depth = np.where(hit, t, 0.0) + np.where(hit, rng.normal(0, 2e-3, (H, W)), 0.0)
        depth = np.where(hit, depth, 0.0).astype(np.float32)
        rgb = np.clip(127 + 120 * n, 0, 255).astype(np.uint8)
        rgb[~hit] = 0
        frames.append(dict(depth=depth, rgb=np.ascontiguousarray(rgb), extrinsic=c2w))
    return frames


def frames_from_npz(z):
    return [dict(depth=z["depth"][k], rgb=z["rgb"][k], extrinsic=z["extrinsic"][k]) for k in range(len(z["depth"]))]


def reference_filters(depth, extrinsic, scale):
    """tsdf_utils.py:83-86,92 applied to one view: min-depth filter, translation / TSDF_scale, depth_trunc."""
    d = depth.copy()
    d[d < np.float32(MIN_DEPTH_BASELINES * BASELINE)] = 0  # :83 (python float vs float32 array compare)
    e = extrinsic.copy()
    e[:3, 3] /= scale  # :85-86
    return d, e, BASELINE * MAX_DEPTH_BASELINES / scale  # :92


def run_open3d(case):
    import open3d as o3d

    if not o3d.__version__.startswith("0.17.0"):
        raise SystemExit(f"open3d {o3d.__version__} is not the reference's 0.17.0")
    c = CASES[case]
    frames = make_frames(case)
    volume = o3d.pipelines.integration.ScalableTSDFVolume(voxel_length=c["voxel"] / 512, sdf_trunc=c["trunc"],
                                                          color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)
    for f in frames:
        d, e, trunc = reference_filters(f["depth"], f["extrinsic"], c["scale"])
        rgbd = o3d.geometry.RGBDImage.create_from_color_and_depth(o3d.geometry.Image(f["rgb"]), o3d.geometry.Image(d),
                                                                   depth_scale=c["scale"], depth_trunc=trunc,
                                                                   convert_rgb_to_intensity=False)
        intr = o3d.camera.PinholeCameraIntrinsic(c["W"], c["H"], c["fx"], c["fy"], c["cx"], c["cy"])
        volume.integrate(rgbd, intr, np.linalg.inv(e))
    vpc = volume.extract_voxel_point_cloud()
    mesh = volume.extract_triangle_mesh()
    out = dict(voxel_points=np.asarray(vpc.points), voxel_tsdf01=np.asarray(vpc.colors)[:, 0],
               mesh_vertices=np.asarray(mesh.vertices).copy(), mesh_triangles=np.asarray(mesh.triangles).copy(),
               mesh_colors=np.asarray(mesh.vertex_colors).copy())
    mesh.scale(c["scale"], (0, 0, 0))
    mesh.compute_vertex_normals()
    out.update(mesh_vertices_scaled=np.asarray(mesh.vertices), mesh_vertex_normals=np.asarray(mesh.vertex_normals),
               backend=f"open3d {o3d.__version__}")
    return frames, out


def oracle_volume(case, frames=None):
    """The CPU restatement on the same frames (also what the consuming tests run)."""
    from oracle import oracle as orc

    c = CASES[case]
    frames = frames or make_frames(case)
    vol = orc.OracleTSDFVolume(c["voxel"] / 512, c["trunc"], with_color=True)
    for f in frames:
        d, e, trunc = reference_filters(f["depth"], f["extrinsic"], c["scale"])
        vol.integrate(d, f["rgb"], c["W"], c["H"], c["fx"], c["fy"], c["cx"], c["cy"], np.linalg.inv(e), depth_scale=c["scale"],
                      depth_trunc=trunc)
    return vol


def oracle_readback(case, vol):
    """What Open3D's extract_voxel_point_cloud / extract_triangle_mesh return, from the oracle volume."""
    from oracle import oracle as orc

    c = CASES[case]
    vl = c["voxel"] / 512
    units = vol.unit_indices()
    pts, vals = [], []
    ii = np.arange(4096)
    loc = np.stack([ii >> 8, (ii >> 4) & 15, ii & 15], -1)
    for i, idx in enumerate(units):
        t, w, _ = vol.unit_data(i)
        keep = (w != 0) & (t < np.float32(0.98)) & (t >= np.float32(-0.98))
        pts.append((idx[None, :] * 16 + loc[keep] + 0.5) * vl)  # origin + (i + 0.5) * voxel_length, fp64
        vals.append((t[keep].astype(np.float64) + 1.0) * 0.5)
    b0 = units.min(axis=0)
    nb = units.max(axis=0) - b0 + 2
    tw, _, outside = vol.export_bricks(b0, nb)
    assert outside == 0
    col = np.zeros((tw.shape[0], 4096, 3), np.float64)
    for i, idx in enumerate(units):
        b = idx - b0
        col[(b[0] * nb[1] + b[1]) * nb[2] + b[2]] = vol.unit_data(i)[2]
    m = orc.extract_mesh_from_bricks(tw, b0, nb, vl, color=col)
    return dict(voxel_points=np.concatenate(pts) if pts else np.zeros((0, 3)), voxel_tsdf01=np.concatenate(vals) if vals else np.zeros(0),
                mesh_vertices=m["vertices"], mesh_triangles=m["triangles"], mesh_colors=m["colors"],
                mesh_vertices_scaled=m["vertices"] * c["scale"], backend="oracle (plumbing check, NOT golden)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="open3d", choices=["open3d", "oracle"])
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for case in CASES:
        if args.backend == "open3d":
            frames, out = run_open3d(case)
        else:
            frames = make_frames(case)
            out = oracle_readback(case, oracle_volume(case, frames))
        path = os.path.join(args.out, f"tsdf_o3d_{case}.npz")
        np.savez_compressed(path, case=case, depth=np.stack([f["depth"] for f in frames]), rgb=np.stack([f["rgb"] for f in frames]),
                            extrinsic=np.stack([f["extrinsic"] for f in frames]), **out)
        print("wrote", path, out["backend"], "voxels", len(out["voxel_tsdf01"]), "triangles", len(out["mesh_triangles"]))


if __name__ == "__main__":
    main()
