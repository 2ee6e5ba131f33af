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


# ---------------------------------------------------------------------------------------------------------------
# Table kernel (render_table_kernel, the default): the same loop with (a) "done" kept in the SIGN of T and (b) the
# exponent evaluated in the log2 domain from per-column / per-row terms, alpha = 2^e with log2(opacity) folded in.
def table_pixel(recs, px, py, bx0=0.0, by0=0.0):
    """float32 restatement of the table blend (blend_stage_alpha + blend_stage_apply) for one pixel: e = fma(v, dy, u') + w; skip if e > thr;
    alpha = min(0.99, 2^e); three chained predicates; T keeps |T| with the sign bit set once saturated."""
    LOG2E = F(1.4426950408889634)
    T, C, D, n = F(1), np.zeros(3, F), F(0), 0
    for (gx, gy, z, op, a, b, c, col) in recs:
        if op == 0:  # the sentinel slot: u' = -1e30 -> alpha 0
            e, thr = F(-1e30), F(0)
        else:
            ap, bp, cp = F(F(-0.72134752044448170) * a), F(-LOG2E * b), F(F(-0.72134752044448170) * c)
            thr = F(np.log2(op))
            dx, dy = F(gx - px), F(gy - py)
            u = F(np.float64(F(ap * dx)) * np.float64(dx) + np.float64(thr))  # fmaf(ap*dx, dx, l2o): one rounding
            v = F(bp * dx)
            w = F(F(cp * dy) * dy)
            e = F(F(np.float64(v) * np.float64(dy) + np.float64(u)) + w)      # FFMA2 then FADD2
        alpha = min(F(0.99), F(np.exp2(e)))
        oma = F(1 - alpha)
        tt = F(T * oma)
        p1 = not (e > thr)
        p2 = p1 and not (alpha < F(1.0 / 255.0))
        ok = p2 and not (tt < F(0.0001))
        sat = p2 and (tt < F(0.0001))
        if ok:
            ww = F(alpha * T)
            C = (col * ww + C).astype(F)
            D = F(z * ww + D)
            T = F(T * oma)
            n += 1
        if sat:
            T = -abs(T)
    return abs(T), C, D, n, bool(T < 0)


def test_sign_of_T_is_an_exact_done_flag():
    """With the reference's own alpha (same power expression) the sign-bit scheme reproduces the reference loop bit for bit:
    after saturation T < 0, so T*(1-alpha) < 1e-4 for every later record and nothing sticks; |T| is the reference's T."""
    rng = np.random.default_rng(15)

    def sign_pixel(recs, px, py):
        T, C, D, n = F(1), np.zeros(3, F), F(0), 0
        for (gx, gy, z, op, a, b, c, col) in recs:
            dx, dy = F(gx - px), F(gy - py)
            power = F(F(-0.5) * F(F(a * dx * dx) + F(c * dy * dy)) - F(b * dx * dy))
            alpha = min(F(0.99), F(op * np.exp(power)))
            tt = F(T * F(1 - alpha))
            p2 = (not (power > 0)) and not (alpha < F(1.0 / 255.0))
            ok, sat = p2 and not (tt < F(0.0001)), p2 and (tt < F(0.0001))
            if ok:
                C = (col * F(alpha * T) + C).astype(F)
                D = F(z * F(alpha * T) + D)
                T = tt
                n += 1
            if sat:
                T = -abs(T)
        return abs(T), C, D, n

    sat = 0
    for trial in range(300):
        recs = _records(rng, int(rng.integers(0, 60)), dense=trial % 2 == 1)
        px, py = F(rng.integers(0, 16)), F(rng.integers(0, 16))
        rT, rC, rD, rn = reference_pixel(recs, px, py)
        sT, sC, sD, sn = sign_pixel(recs + [SENTINEL], px, py)
        assert sT == rT and sn == rn
        np.testing.assert_allclose(sC, rC, rtol=0, atol=2e-6)
        sat += rT < F(0.001)
    assert sat > 50


def test_two_stage_blend_with_inert_records_equals_the_sign_scheme():
    """The pipelined loop of render_table_kernel splits a record into a T-independent stage (alpha, forced to 0 where the
    reference `continue`s) and the serial stage (T*(1-alpha) >= 1e-4 ? accumulate, T = test_T : set the sign) that no longer
    looks at the `continue` predicates: a record with alpha 0 is inert (T*1 = T >= 1e-4 while the pixel lives, negative once
    it is done; colour += c*0), so transmittance, contributor count and colour are those of the reference loop."""
    rng = np.random.default_rng(35)

    def two_stage_pixel(recs, px, py):
        T, C, D, n = F(1), np.zeros(3, F), F(0), 0
        for (gx, gy, z, op, a, b, c, col) in recs:
            dx, dy = F(gx - px), F(gy - py)
            power = F(F(-0.5) * F(F(a * dx * dx) + F(c * dy * dy)) - F(b * dx * dy))
            alpha = min(F(0.99), F(op * np.exp(power)))  # stage A
            if (power > 0) or (alpha < F(1.0 / 255.0)):
                alpha = F(0)
            tt = F(T * F(1 - alpha))  # stage B
            if tt >= F(0.0001):
                w = F(alpha * T)
                C = (col * w + C).astype(F)
                D = F(z * w + D)
                T = tt
                n += int(alpha > 0)
            else:
                T = -abs(T)
        return abs(T), C, D, n

    sat = 0
    for trial in range(300):
        recs = _records(rng, int(rng.integers(0, 60)), dense=trial % 2 == 1)
        px, py = F(rng.integers(0, 16)), F(rng.integers(0, 16))
        rT, rC, rD, rn = reference_pixel(recs, px, py)
        sT, sC, sD, sn = two_stage_pixel(recs + [SENTINEL, SENTINEL, SENTINEL], px, py)
        assert sT == rT and sn == rn
        np.testing.assert_allclose(sC, rC, rtol=0, atol=2e-6)
        np.testing.assert_allclose(sD, rD, rtol=0, atol=1e-5)
        sat += rT < F(0.001)
    assert sat > 50


def test_table_exponent_tracks_the_reference_within_the_parity_budget():
    """The log2-domain exponent differs from the reference's expression by rounding only: alpha within ~1e-5 relative (far
    inside north_star's 1e-4), so pixels agree unless a threshold sits inside that rounding band -- counted, and rare."""
    rng = np.random.default_rng(25)
    flips, worst, total = 0, 0.0, 0
    for trial in range(400):
        recs = _records(rng, int(rng.integers(1, 60)), dense=trial % 2 == 1)
        px, py = F(rng.integers(0, 16)), F(rng.integers(0, 16))
        rT, rC, rD, rn = reference_pixel(recs, px, py)
        tT, tC, tD, tn, done = table_pixel(recs + [SENTINEL], px, py)
        total += 1
        if tn != rn:  # a record sat within rounding of a threshold: different contributor set
            flips += 1
            continue
        worst = max(worst, float(np.abs(tC - rC).max()), abs(float(tT) - float(rT)))
    assert flips <= 0.02 * total, (flips, total)
    assert worst < 2e-5, worst
