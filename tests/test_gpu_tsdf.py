"""GPU parity tests of the TSDF half: CUDA brick volume (through the C ABI) vs the CPU oracle
(Open3D-0.17 restatement) on the same seeded depth frames.

Bar: (tsdf, weight) BIT-EXACT against the oracle (same fp32 operation order, no FMA), which is
far inside the north_star's 1e-3; colour within 1e-4 (the oracle keeps colour in fp64 like
Open3D, the GPU in fp32).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 160, 120
FX = FY = 150.0
CX, CY = 80.0, 60.0
VL = 2.0 / 128  # 128^3 lattice over [-1,1]^3
TRUNC = 0.05
B0, NB = (-4, -4, -4), (8, 8, 8)


def _look_at(pos, target=(0, 0, 0)):
    """world->camera 4x4 (OpenCV axes) of a camera at `pos` looking at `target`."""
    pos = np.asarray(pos, float)
    f = np.asarray(target, float) - pos
    f /= np.linalg.norm(f)
    r = np.cross(f, [0, 0, 1.0])
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    c2w = np.eye(4)
    c2w[:3, :3] = np.stack([r, d, f], axis=1)
    c2w[:3, 3] = pos
    return np.linalg.inv(c2w)


def _sphere_depth(w2c, radius=0.6):
    """Analytic depth map of a sphere at the origin + a ragged hole and some zero pixels."""
    c2w = np.linalg.inv(w2c)
    ys, xs = np.mgrid[0:H, 0:W]
    dirs = np.stack([(xs - CX) / FX, (ys - CY) / FY, np.ones_like(xs, float)], -1)
    dw = dirs @ c2w[:3, :3].T
    o = c2w[:3, 3]
    a = (dw * dw).sum(-1)
    b = 2 * (dw @ o)
    c = o @ o - radius ** 2
    disc = b * b - 4 * a * c
    t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 0.0)
    depth = np.where(disc > 0, t, 0.0).astype(np.float32)  # z-depth since dirs has z = 1
    depth[10:20, 30:50] = 0
    return depth


def _views(n=5, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        az = 2 * np.pi * k / n + rng.uniform(-0.1, 0.1)
        pos = 2.2 * np.array([np.cos(az), np.sin(az), 0.3 * np.sin(3 * az)])
        w2c = _look_at(pos)
        depth = _sphere_depth(w2c)
        rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        out.append((depth, rgb, w2c))
    return out


def _gpu_volume(dev, with_color=True):
    from gs2mesh_b200.tsdf import TSDFVolume

    return TSDFVolume(VL, TRUNC, B0, NB, with_color=with_color, device=dev)


def test_brick_volume_bit_exact_vs_oracle(oracle, gsb_lib, cuda_device):
    import torch

    views = _views(5)
    ovol = oracle.OracleTSDFVolume(VL, TRUNC, with_color=True)
    gvol = _gpu_volume(cuda_device)
    for depth, rgb, w2c in views:
        n_units = ovol.integrate(depth, rgb, W, H, FX, FY, CX, CY, w2c, depth_scale=1.0, depth_trunc=4.0)
        prepared = gvol.prepare_depth(depth, W, H, depth_scale=1.0, depth_trunc=4.0)
        gvol.integrate(prepared, rgb, W, H, FX, FY, CX, CY, w2c)
        touched, outside, _ = gvol.last_stats()
        assert outside == 0
        assert touched == n_units  # same set of volume units opened per frame
    torch.cuda.synchronize()
    tw, alloc, n_out = ovol.export_bricks(B0, NB)
    assert n_out == 0 and alloc.sum() == ovol.num_units
    got = gvol.bricks().cpu().numpy()
    np.testing.assert_array_equal(got[..., 1], tw[..., 1])  # weights
    np.testing.assert_array_equal(got[..., 0], tw[..., 0])  # tsdf, bit-exact
    assert (got[alloc == 0] == 0).all()  # bricks Open3D would not allocate stay untouched
    assert (tw[..., 1] > 0).sum() > 10000
    # unit by unit, keyed by lattice index; colour: fp32 running mean vs Open3D's fp64
    from tests.volume_compare import assert_units_equal

    assert assert_units_equal(gvol, ovol) == ovol.num_units == gvol.num_bricks()


def test_prepare_depth_matches_reference_filters(oracle, gsb_lib, cuda_device):
    """tsdf_utils.py:78-93 order: masks, min-depth, /scale, >= trunc -> 0."""
    rng = np.random.default_rng(1)
    depth = rng.uniform(0.0, 5.0, (H, W)).astype(np.float32)
    mask = rng.integers(0, 2, (H, W)).astype(np.uint8)
    gvol = _gpu_volume(cuda_device, with_color=False)
    got = gvol.prepare_depth(depth, W, H, mask=mask, min_depth=0.7, depth_scale=2.0, depth_trunc=1.5).cpu().numpy()
    d = depth * mask
    d = np.where(d < np.float32(0.7), 0, d).astype(np.float32)
    want = oracle.depth_convert(d, 2.0, 1.5)
    np.testing.assert_array_equal(got, want)
    # expected-depth normalisation from the rasterizer's (sum z*alpha*T, final_T)
    T = rng.uniform(0, 1, (H, W)).astype(np.float32)
    got = gvol.prepare_depth(depth, W, H, final_T=T, alpha_min=0.5).cpu().numpy()
    alpha = np.float32(1) - T
    want = np.where(alpha > 0.5, depth / alpha, 0).astype(np.float32)
    np.testing.assert_array_equal(got, want)


def test_weight_counts_and_idempotent_mean(gsb_lib, cuda_device):
    depth, rgb, w2c = _views(1)[0]
    gvol = _gpu_volume(cuda_device)
    for _ in range(4):
        gvol.integrate(gvol.prepare_depth(depth, W, H), rgb, W, H, FX, FY, CX, CY, w2c)
    b4 = gvol.bricks().cpu().numpy().copy()
    g1 = _gpu_volume(cuda_device)
    g1.integrate(g1.prepare_depth(depth, W, H), rgb, W, H, FX, FY, CX, CY, w2c)
    b1 = g1.bricks().cpu().numpy()
    np.testing.assert_array_equal(b4[..., 1], 4 * b1[..., 1])
    np.testing.assert_allclose(b4[..., 0], b1[..., 0], atol=1e-6)


def test_view_order_independence_and_sum_form_roundtrip(gsb_lib, cuda_device):
    views = _views(6, seed=3)

    def fuse(order):
        v = _gpu_volume(cuda_device, with_color=False)
        for i in order:
            v.integrate(v.prepare_depth(views[i][0], W, H), None, W, H, FX, FY, CX, CY, views[i][2])
        return v

    a = fuse(range(6))
    b = fuse([3, 0, 5, 1, 4, 2])
    ta, tb = a.bricks().cpu().numpy(), b.bricks().cpu().numpy()
    np.testing.assert_array_equal(ta[..., 1], tb[..., 1])
    np.testing.assert_allclose(ta[..., 0], tb[..., 0], atol=2e-6)
    # sharded merge: (mean,w) -> (sum,w), add, -> (mean,w) equals the sequential volume
    # (the two shards open different bricks in different pool slots: add them brick by brick through the view window)
    s0, s1 = fuse([0, 2, 4]), fuse([1, 3, 5])
    s0.to_sums()
    s1.to_sums()
    sums = s0.bricks().cpu().numpy() + s1.bricks().cpu().numpy()
    wsum = sums[..., 1]
    tm = np.where(wsum > 0, sums[..., 0] / np.maximum(wsum, 1), 0).astype(np.float32)
    np.testing.assert_array_equal(wsum, ta[..., 1])
    np.testing.assert_allclose(tm, ta[..., 0], atol=2e-6)
    s0.from_sums()  # and back: (sum, w) -> (mean, w) restores the shard
    np.testing.assert_allclose(s0.bricks().cpu().numpy(), fuse([0, 2, 4]).bricks().cpu().numpy(), atol=1e-6)


def test_dense_export_layout(gsb_lib, cuda_device):
    depth, rgb, w2c = _views(1)[0]
    gvol = _gpu_volume(cuda_device, with_color=False)
    gvol.integrate(gvol.prepare_depth(depth, W, H), None, W, H, FX, FY, CX, CY, w2c)
    tsdf, weight = gvol.dense()
    bricks = gvol.bricks().cpu().numpy().reshape(NB[0], NB[1], NB[2], 16, 16, 16, 2)
    dense = np.transpose(bricks, (0, 3, 1, 4, 2, 5, 6)).reshape(128, 128, 128, 2)
    np.testing.assert_array_equal(tsdf.cpu().numpy(), dense[..., 0])
    np.testing.assert_array_equal(weight.cpu().numpy(), dense[..., 1])


def test_pool_exhaustion_is_counted_and_grown(gsb_lib, cuda_device):
    """The volume is unbounded (a hash of bricks, like Open3D's); what is finite is the brick pool.  Bricks that do not fit
    are counted, never written out of bounds; ensure_capacity() doubles the pool and reports that the batch must be fused
    again.  The regrown volume equals one that had room from the start, and grow() keeps what was already fused."""
    from gs2mesh_b200.tsdf import TSDFVolume

    views = _views(3, seed=5)
    big = _gpu_volume(cuda_device)
    for depth, rgb, w2c in views:
        big.integrate(big.prepare_depth(depth, W, H), rgb, W, H, FX, FY, CX, CY, w2c)
    assert big.pool_stats()["dropped_total"] == 0
    small = TSDFVolume(VL, TRUNC, B0, NB, with_color=True, device=cuda_device, pool_bricks=8)
    attempts = 0
    for attempts in range(1, 16):
        for depth, rgb, w2c in views:
            small.integrate(small.prepare_depth(depth, W, H), rgb, W, H, FX, FY, CX, CY, w2c)
        if attempts == 1:
            st = small.pool_stats()
            assert st["dropped"] > 0 and st["dropped_total"] >= st["dropped"] and st["bricks"] <= 8
        if small.ensure_capacity():
            break
        small.reset()
    assert attempts > 1 and small.pool_bricks >= big.num_bricks() and small.num_bricks() == big.num_bricks()
    np.testing.assert_array_equal(small.bricks().cpu().numpy(), big.bricks().cpu().numpy())
    a, b = small.export_units(), big.export_units()
    assert set(a) == set(b)
    for k in a:
        np.testing.assert_array_equal(a[k][1], b[k][1])
    # grow() re-houses the open bricks
    before = small.export_units()
    small.grow()
    after = small.export_units()
    assert set(before) == set(after)
    for k in before:
        np.testing.assert_array_equal(before[k][0], after[k][0])
        np.testing.assert_array_equal(before[k][1], after[k][1])
    depth, rgb, w2c = views[0]
    small.integrate(small.prepare_depth(depth, W, H), rgb, W, H, FX, FY, CX, CY, w2c)  # and stays usable
    big.integrate(big.prepare_depth(depth, W, H), rgb, W, H, FX, FY, CX, CY, w2c)
    np.testing.assert_array_equal(small.bricks().cpu().numpy(), big.bricks().cpu().numpy())


def test_unbounded_volume_far_from_the_origin(oracle, gsb_lib, cuda_device):
    """Reference semantics of tsdf_utils.py:51-56,85-93: ScalableTSDFVolume has no origin and no extent.  A scene centred at
    (5, -3, 2) fused with TSDF_scale = 0.1 (translations / 0.1, depth_scale 0.1: lattice indices around (50, -30, 20) * 16)
    opens exactly the oracle's units -- none dropped, none clipped -- with bit-identical voxels; likewise an off-centre
    principal point.  Frames = the seeded ones of tests/golden/make_tsdf_golden.py."""
    from gs2mesh_b200.tsdf import TSDFVolume
    from tests.golden import make_tsdf_golden as g
    from tests.volume_compare import assert_units_equal

    for case in ("scaled_shifted", "offcentre"):
        c = g.CASES[case]
        frames = g.make_frames(case)
        ovol = g.oracle_volume(case, frames)
        vol = TSDFVolume(c["voxel"] / 512, c["trunc"], with_color=True, device=cuda_device, pool_bricks=4096)
        for f in frames:
            d, e, trunc = g.reference_filters(f["depth"], f["extrinsic"], c["scale"])
            prepared = vol.prepare_depth(d, c["W"], c["H"], depth_scale=c["scale"], depth_trunc=trunc)
            vol.integrate(prepared, f["rgb"], c["W"], c["H"], c["fx"], c["fy"], c["cx"], c["cy"], np.linalg.inv(e))
        st = vol.pool_stats()
        assert st["dropped_total"] == 0
        n = assert_units_equal(vol, ovol)
        assert n == st["bricks"] > 50
        if case == "scaled_shifted":
            assert np.abs(vol.brick_indices().cpu().numpy()).max() > 40  # nowhere near a window around the origin


def test_invalid_inputs(gsb_lib, cuda_device):
    gvol = _gpu_volume(cuda_device)
    depth = np.zeros((H, W), np.float32)
    with pytest.raises(RuntimeError, match="Unsupported image format"):
        gvol.integrate(depth[:-1], None, W, H, FX, FY, CX, CY, np.eye(4))
    with pytest.raises(RuntimeError, match="singular"):
        gvol.integrate(depth, None, W, H, FX, FY, CX, CY, np.zeros((4, 4)))


def test_mask_filters_bit_exact(oracle, gsb_lib, cuda_device):
    """gsb_mask_morphology / filter_object_mask == the cv2 restatement (tsdf_utils.py:69-77), ragged sizes and kernel sizes."""
    import torch

    from gs2mesh_b200 import _lib
    from gs2mesh_b200.tsdf import filter_object_mask
    from tests.test_oracle_tsdf import _blob_mask

    rng = np.random.default_rng(11)
    for (h, w), k in (((37, 53), 10), ((130, 67), 3), ((20, 200), 7), ((9, 9), 10), ((31, 17), 1), ((240, 320), 64), ((1, 1), 2)):
        m = _blob_mask(rng, h, w).astype(np.uint8) * rng.integers(1, 255, (h, w), dtype=np.uint8)  # general u8, not only 0/1
        src = torch.as_tensor(m).to(cuda_device)
        dst = torch.empty_like(src)
        with torch.cuda.device(cuda_device):
            stream = torch.cuda.current_stream().cuda_stream
            for dilate in (0, 1):
                _lib.check(gsb_lib.gsb_mask_morphology(_lib.ptr(src), w, h, k, dilate, _lib.ptr(dst), stream))
                np.testing.assert_array_equal(dst.cpu().numpy(), oracle.mask_morphology(m, k, bool(dilate)))
        got = filter_object_mask(m > 0, k, max(1, k // 2), invert=bool(k & 1), device=cuda_device)
        np.testing.assert_array_equal(got.cpu().numpy() > 0, oracle.filter_object_mask(m > 0, k, max(1, k // 2), invert=bool(k & 1)))
    try:
        import cv2
    except ImportError:
        cv2 = None
    if cv2 is not None:  # and against OpenCV itself at the reference's default sizes
        m = _blob_mask(rng, 480, 640)
        closing = cv2.morphologyEx(m.astype(np.uint8), cv2.MORPH_CLOSE, np.ones((10, 10), np.uint8))
        want = cv2.erode(closing, np.ones((10, 10), np.uint8), iterations=1) > 0.5
        np.testing.assert_array_equal(filter_object_mask(m, 10, 10, device=cuda_device).cpu().numpy() > 0, want)
    src = torch.zeros(8, 8, dtype=torch.uint8, device=cuda_device)
    assert gsb_lib.gsb_mask_morphology(_lib.ptr(src), 8, 8, 65, 0, _lib.ptr(torch.empty_like(src)), None) == _lib.GSB_ERR_INVALID
    assert gsb_lib.gsb_mask_morphology(_lib.ptr(src), 8, 8, 3, 0, _lib.ptr(src), None) == _lib.GSB_ERR_INVALID
