"""Property test of the exact (Gaussian, tile) culling rule of gs2mesh_b200/csrc/gsb_raster.cu (make_footprint,
rect_can_contribute, shrink_rect), restated in float32 numpy: over random conics, opacities and positions, every tile
that holds a pixel the reference blend loop would NOT skip (forward.cu:336-346: power <= 0 and
opacity * exp(power) >= 1/255) is kept -- by the tile test and by the tightened candidate rectangle.  That is the
property the "bit-identical to rectangle binning" claim rests on; the GPU tests check it on rendered scenes, this checks
the arithmetic (margins included) far outside them: needle-shaped, huge, tiny, nearly transparent Gaussians."""
import numpy as np

F = np.float32
TILE = 16


def make_footprint(a, b, c, two_tau):
    degenerate = not (a > 0 and c > 0 and F(a * c) - F(b * b) > 0)
    nb_a = F(0) if degenerate else F(-b / a)
    nb_c = F(0) if degenerate else F(-b / c)
    return dict(a=F(a), b=F(b), c=F(c), nb_a=nb_a, nb_c=nb_c, two_tau=F(two_tau), degenerate=degenerate)


def rect_can_contribute(gx, gy, f, x0, y0, x1, y1):
    if f["degenerate"]:
        return True
    a, b, c = f["a"], f["b"], f["c"]
    dx_lo, dx_hi = F(gx - x1), F(gx - x0)
    dy_lo, dy_hi = F(gy - y1), F(gy - y0)
    if dx_lo <= 0 <= dx_hi and dy_lo <= 0 <= dy_hi:
        return True
    DX, DY = max(abs(dx_lo), abs(dx_hi)), max(abs(dy_lo), abs(dy_hi))
    margin = F(1e-3) + F(8e-6) * F(a * DX * DX + c * DY * DY + F(2) * abs(b) * DX * DY)
    dx = dx_lo if abs(dx_lo) < abs(dx_hi) else dx_hi
    dy = min(dy_hi, max(dy_lo, F(f["nb_c"] * dx)))
    qv = F(a * dx * dx + F(2) * b * dx * dy + c * dy * dy)
    ey = dy_lo if abs(dy_lo) < abs(dy_hi) else dy_hi
    ex = min(dx_hi, max(dx_lo, F(f["nb_a"] * ey)))
    qh = F(a * ex * ex + F(2) * b * ex * ey + c * ey * ey)
    return min(qv, qh) <= f["two_tau"] + margin


def tile_rect(px, py, radius, gx, gy):
    x0 = min(gx, max(0, int((px - radius) / TILE)))
    y0 = min(gy, max(0, int((py - radius) / TILE)))
    x1 = min(gx, max(0, int((px + radius + TILE - 1) / TILE)))
    y1 = min(gy, max(0, int((py + radius + TILE - 1) / TILE)))
    return x0, y0, x1, y1


def shrink_rect(r, px, py, f):
    x0, y0, x1, y1 = r
    if f["degenerate"]:
        return r
    ac = F(f["a"] * f["c"])
    det = F(ac - F(f["b"] * f["b"]))
    k = F(ac / det)
    if not k <= 1000:
        return r
    tau = F(F(f["two_tau"] + F(2e-3)) * F(F(1) + F(F(5e-5) * k)))
    if not tau > 0:
        return x0, y0, x0, y0
    hx = F(F(np.sqrt(F(F(tau * f["c"]) / det))) * F(1.0001) + F(0.01))
    hy = F(F(np.sqrt(F(F(tau * f["a"]) / det))) * F(1.0001) + F(0.01))
    inv = F(1.0 / TILE)
    lox, hix = np.floor(F(F(px - hx) * inv)), np.floor(F(F(px + hx) * inv))
    loy, hiy = np.floor(F(F(py - hy) * inv)), np.floor(F(F(py + hy) * inv))
    nx0, nx1 = max(float(x0), lox), min(float(x1), hix + 1)
    ny0, ny1 = max(float(y0), loy), min(float(y1), hiy + 1)
    if not (nx0 < nx1 and ny0 < ny1):
        return x0, y0, x0, y0
    return int(nx0), int(ny0), int(nx1), int(ny1)


def blended_somewhere(px, py, a, b, c, opacity, tx, ty, box=None):
    """forward.cu:336-346 over the 256 pixels of tile (tx, ty) -- or over the pixel box (x0, y0, x1, y1) -- in float32
    like the kernel."""
    if box is None:
        box = (tx * TILE, ty * TILE, tx * TILE + TILE - 1, ty * TILE + TILE - 1)
    xs = np.arange(box[0], box[2] + 1).astype(F)[None, :]
    ys = np.arange(box[1], box[3] + 1).astype(F)[:, None]
    dx, dy = F(px) - xs, F(py) - ys
    power = F(-0.5) * (F(a) * dx * dx + F(c) * dy * dy) - F(b) * dx * dy
    alpha = np.minimum(F(0.99), F(opacity) * np.exp(power.astype(F)))
    return bool(((power <= 0) & (alpha >= F(1.0 / 255.0))).any())


def random_gaussian(rng):
    """2-D covariance of forward.cu:74-113 (+0.3 low-pass) -> conic, 3-sigma radius, opacity, centre."""
    s1 = np.exp(rng.uniform(np.log(0.05), np.log(60.0)))
    s2 = s1 * np.exp(rng.uniform(np.log(0.01), 0.0))  # anisotropy up to 100:1
    th = rng.uniform(0, np.pi)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    cov = R @ np.diag([s1 * s1, s2 * s2]) @ R.T
    cxx, cxy, cyy = F(cov[0, 0] + 0.3), F(cov[0, 1]), F(cov[1, 1] + 0.3)
    det = F(cxx * cyy - cxy * cxy)
    a, b, c = F(cyy / det), F(-cxy / det), F(cxx / det)
    mid = F(0.5) * (cxx + cyy)
    lam = mid + F(np.sqrt(max(F(0.1), F(mid * mid - det))))
    radius = int(np.ceil(F(3.0) * np.sqrt(lam)))
    opacity = F(rng.choice([rng.uniform(0.004, 0.02), rng.uniform(0.02, 1.0), 1.0], p=[0.3, 0.6, 0.1]))
    px, py = F(rng.uniform(-40, 360)), F(rng.uniform(-40, 280))
    return px, py, a, b, c, opacity, radius


def test_no_contributing_tile_is_ever_dropped():
    rng = np.random.default_rng(20260923)
    gx, gy = 20, 15  # 320 x 240 pixels
    checked = kept = contributing = 0
    for _ in range(1500):
        px, py, a, b, c, opacity, radius = random_gaussian(rng)
        ref = tile_rect(px, py, radius, gx, gy)
        if (ref[2] - ref[0]) * (ref[3] - ref[1]) == 0:
            continue
        f = make_footprint(a, b, c, F(2.0) * F(np.log(F(255.0) * opacity)))
        cand = shrink_rect(ref, px, py, f)
        for ty in range(ref[1], ref[3]):
            for tx in range(ref[0], ref[2]):
                in_cand = cand[0] <= tx < cand[2] and cand[1] <= ty < cand[3]
                x0, y0 = F(tx * TILE), F(ty * TILE)
                keep = in_cand and rect_can_contribute(px, py, f, x0, y0, F(x0 + TILE - 1), F(y0 + TILE - 1))
                contrib = blended_somewhere(px, py, a, b, c, opacity, tx, ty)
                checked += 1
                kept += keep
                contributing += contrib
                assert keep or not contrib, dict(px=px, py=py, conic=(a, b, c), opacity=opacity, tile=(tx, ty), ref=ref, cand=cand)
    # the rule is also worth having: it removes a good part of the rectangle's tiles and keeps few that contribute nothing
    assert checked > 20_000 and contributing > 2_000
    assert kept < 0.8 * checked and kept <= 1.6 * contributing + 50, (checked, kept, contributing)


def test_reference_rectangle_always_covers_the_footprint():
    """The starting point: the reference's own 3-sigma rectangle (auxiliary.h:46-56) is what forward.cu walks, so a tile
    outside it is never blended by the reference either -- nothing outside it needs a test."""
    rng = np.random.default_rng(7)
    for _ in range(300):
        px, py, a, b, c, opacity, radius = random_gaussian(rng)
        ref = tile_rect(px, py, radius, 20, 15)
        f = make_footprint(a, b, c, F(2.0) * F(np.log(F(255.0) * opacity)))
        cand = shrink_rect(ref, px, py, f)
        assert ref[0] <= cand[0] <= cand[2] <= ref[2] and ref[1] <= cand[1] <= cand[3] <= ref[3]


def test_per_warp_block_test_of_the_blend_kernels_is_conservative():
    """The blend kernels apply the same rule to a warp's 8x8 (dual) or 8x4 (compact) pixel block, with
    two_tau = 2 ln(255 opacity) + 1e-3 from the fast logarithm: a record the block test rejects has no pixel in the block
    the reference loop would blend."""
    rng = np.random.default_rng(11)
    rejected = hit = 0
    for _ in range(1200):
        px, py, a, b, c, opacity, radius = random_gaussian(rng)
        f = make_footprint(a, b, c, F(F(2.0) * F(np.log(F(255.0) * opacity)) + F(1e-3)))
        ref = tile_rect(px, py, radius, 20, 15)
        for ty in range(ref[1], ref[3]):
            for tx in range(ref[0], ref[2]):
                for bw, bh in ((8, 8), (8, 4)):
                    for by in range(0, TILE, bh):
                        for bx in range(0, TILE, bw):
                            x0, y0 = tx * TILE + bx, ty * TILE + by
                            keep = rect_can_contribute(px, py, f, F(x0), F(y0), F(x0 + bw - 1), F(y0 + bh - 1))
                            if keep:
                                hit += 1
                                continue
                            rejected += 1
                            assert not blended_somewhere(px, py, a, b, c, opacity, tx, ty, box=(x0, y0, x0 + bw - 1, y0 + bh - 1)), \
                                dict(px=px, py=py, conic=(a, b, c), opacity=opacity, block=(x0, y0, bw, bh))
    assert rejected > 50_000 and hit > 10_000
