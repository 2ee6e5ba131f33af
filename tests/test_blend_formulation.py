"""The branch-free blend of render_compact_kernel (gs2mesh_b200/csrc/gsb_raster.cu) against the reference loop
(forward.cu:325-362), both restated in float32 numpy for one pixel: the three thresholds (`power > 0`, `alpha < 1/255`,
`T(1-alpha) < 1e-4` -> done) become predicates gating a weight w = alpha*T and the transmittance update.  Same
transmittance bit for bit, same contributor decisions; colour differs only by fma(c, alpha*T, C) vs C + c*alpha*T.
Also: the all-zero sentinel record that pads an odd hit list is inert, and records after saturation never contribute."""
import numpy as np

F = np.float32


def reference_pixel(recs, px, py):
    T, C, D, done, n = F(1), np.zeros(3, F), F(0), False, 0
    for (gx, gy, z, op, a, b, c, col) in recs:
        if done:
            break
        dx, dy = F(gx - px), F(gy - py)
        power = F(F(-0.5) * F(F(a * dx * dx) + F(c * dy * dy)) - F(b * dx * dy))
        if power > 0:
            continue
        alpha = min(F(0.99), F(op * np.exp(power)))
        if alpha < F(1.0 / 255.0):
            continue
        test_T = F(T * F(1 - alpha))
        if test_T < F(0.0001):
            done = True
            continue
        C = (C + col * alpha * T).astype(F)
        D = F(D + z * alpha * T)
        T = test_T
        n += 1
    return T, C, D, n


def predicated_pixel(recs, px, py):
    T, C, D, done, n = F(1), np.zeros(3, F), F(0), False, 0
    for (gx, gy, z, op, a, b, c, col) in recs:  # every record is evaluated; predicates decide what sticks
        dx, dy = F(gx - px), F(gy - py)
        power = F(F(-0.5) * F(F(a * dx * dx) + F(c * dy * dy)) - F(b * dx * dy))
        alpha = min(F(0.99), F(op * np.exp(power)))
        test_T = F(T * F(1 - alpha))
        cand = (not done) and not (power > 0) and not (alpha < F(1.0 / 255.0))
        sat = cand and test_T < F(0.0001)
        ok = cand and not sat
        done = done or sat
        w = F(alpha * T) if ok else F(0)
        C = (col * w + C).astype(F)
        D = F(z * w + D)
        T = test_T if ok else T
        n += ok
    return T, C, D, n


def _records(rng, n, dense=False):
    recs = []
    for _ in range(n):
        s = np.exp(rng.uniform(np.log(4.0 if dense else 0.5), np.log(12.0)))
        a = c = F(1.0 / (s * s + 0.3))
        b = F(rng.uniform(-0.5, 0.5) * a)
        op = rng.uniform(0.5, 1.0) if dense else rng.choice([rng.uniform(0.003, 0.05), rng.uniform(0.05, 1.0)])
        recs.append((F(rng.uniform(-6, 22)), F(rng.uniform(-6, 22)), F(rng.uniform(1.0, 4.0)), F(op), a, b, c,
                     rng.uniform(0, 1.5, 3).astype(F)))
    return recs


SENTINEL = (F(0), F(0), F(0), F(0), F(0), F(0), F(0), np.zeros(3, F))


def test_predicated_blend_equals_reference_loop():
    rng = np.random.default_rng(5)
    saturated = 0
    for trial in range(300):
        recs = _records(rng, int(rng.integers(0, 60)), dense=trial % 2 == 1)
        px, py = F(rng.integers(0, 16)), F(rng.integers(0, 16))
        rT, rC, rD, rn = reference_pixel(recs, px, py)
        # odd-length lists are padded with the sentinel, and a sentinel anywhere changes nothing
        padded = list(recs) + [SENTINEL]
        if recs:
            padded.insert(int(rng.integers(0, len(recs))), SENTINEL)
        for variant in (recs, padded):
            pT, pC, pD, pn = predicated_pixel(variant, px, py)
            assert pT == rT and pn == rn  # bit-identical transmittance, same contributors
            np.testing.assert_allclose(pC, rC, rtol=0, atol=2e-6)
            np.testing.assert_allclose(pD, rD, rtol=2e-6, atol=1e-6)
        saturated += rT < F(0.001)
    assert saturated > 50  # the done path is exercised


def test_nothing_contributes_after_saturation():
    rng = np.random.default_rng(6)
    opaque = [(F(8), F(8), F(2), F(1.0), F(0.01), F(0), F(0.01), np.ones(3, F))] * 12  # alpha = 0.99 each: T -> 1e-24 would follow
    tail = _records(rng, 30)
    a = predicated_pixel(opaque + tail, F(8), F(8))
    b = predicated_pixel(opaque, F(8), F(8))
    assert a[0] == b[0] and a[3] == b[3] and np.array_equal(a[1], b[1])
    assert a[0] >= F(0.0001)  # the reference stops BEFORE T drops under 1e-4
