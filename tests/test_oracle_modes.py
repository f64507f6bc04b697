"""Unit tests of the oracle itself: RNG known answers, the f32 math primitives of the arithmetic contract,
and how the F32 (kernel-mirror) and F64 (physics) modes relate."""
import ctypes as C

import numpy as np


def test_philox4x32_10_known_answers(oracle):
    """Random123 kat_vectors for philox4x32-10 (Salmon et al. 2011)."""
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox4x32_7_known_answers_and_round_switch(oracle):
    """Random123 kat_vectors for philox4x32-7 (the generator of the library's CONTRACT_FAST7), and the process-wide switch
    that makes orc_draws / orc_trace use it."""
    assert oracle.philox([0, 0, 0, 0], [0, 0], rounds=7) == [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2, rounds=7) == [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], rounds=7) == \
        [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]
    d10 = oracle.draws(4357, 12345, 3)
    oracle.set_philox_rounds(7)
    try:
        d7 = oracle.draws(4357, 12345, 3)
        w = oracle.philox([12345, 0, 3, 0], [4357, 0], rounds=7)
        assert d7[0] == np.float32((w[0] >> 8) * 2.0 ** -24) and d7[2] == np.float32((w[2] >> 12) * 2.0 ** -20)
        # the physics does not care: same escape fraction and hit count within sampling error
        _, s7 = oracle.trace(oracle.scene(theta_max=170.0), oracle.source(), 150_000, seed=8, prec=oracle.F64)
    finally:
        oracle.set_philox_rounds(10)
    assert not np.array_equal(d7, d10) and np.array_equal(oracle.draws(4357, 12345, 3), d10)
    _, s10 = oracle.trace(oracle.scene(theta_max=170.0), oracle.source(), 150_000, seed=8, prec=oracle.F64)
    p7, p10 = s7["n_exit_port"] / 150_000, s10["n_exit_port"] / 150_000
    assert abs(p7 - p10) < 4 * np.sqrt(2 * 0.2445 / 150_000) and s7["n_exit_port"] != s10["n_exit_port"]
    assert abs(s7["n_bounces"] / s10["n_bounces"] - 1) < 4 * np.sqrt(2.0 / 150_000)


def test_f32_primitives_accuracy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    s, c = C.c_float(), C.c_float()
    worst = 0.0
    for u in np.concatenate([rng.random(20000, dtype=np.float32), np.float32([0.0, 0.125, 0.25, 0.5, 0.75, 1 - 2 ** -24])]):
        L.orc_sincos2pi_f32(float(u), C.byref(s), C.byref(c))
        x = 2 * np.pi * float(u)
        worst = max(worst, abs(s.value - np.sin(x)), abs(c.value - np.cos(x)))
    assert worst < 4e-7, worst
    # table forms used for the RNG's fixed-point azimuths: 13-bit lookup, 20-bit lookup + second-order rotation
    worst, norm = 0.0, 0.0
    for q in np.concatenate([rng.integers(0, 1 << 20, 20000), [0, 127, 128, 1 << 18, (1 << 20) - 1]]):
        L.orc_sincos2pi_q20(int(q), C.byref(s), C.byref(c))
        x = 2 * np.pi * int(q) / (1 << 20)
        worst = max(worst, abs(s.value - np.sin(x)), abs(c.value - np.cos(x)))
        norm = max(norm, abs(s.value ** 2 + c.value ** 2 - 1.0))
        if q % 128 == 0:
            s2, c2 = C.c_float(), C.c_float()
            L.orc_sincos2pi_q13(int(q) >> 7, C.byref(s2), C.byref(c2))
            assert (s2.value, c2.value) == (s.value, c.value)
    assert worst < 4e-7 and norm < 4e-7, (worst, norm)
    worst = 0.0
    for x in rng.uniform(-30, 30, 20000).astype(np.float32):
        L.orc_sincos_f32(float(x), C.byref(s), C.byref(c))
        worst = max(worst, abs(s.value - np.sin(float(x))), abs(c.value - np.cos(float(x))))
    assert worst < 2e-6, worst
    worst = 0.0
    for k in np.concatenate([rng.integers(1, (1 << 20) + 1, 20000), [1, 2, 3, 1 << 10, (1 << 19) - 1, 1 << 19, (1 << 20) - 1, 1 << 20]]):
        got = L.orc_log_u20(int(k))
        ref = np.log(int(k) * 2.0 ** -20)
        worst = max(worst, abs(got - ref) / max(1.0, abs(ref)))
    assert worst < 3e-7, worst
    assert L.orc_log_u20(1 << 20) == 0.0


def test_draw_record_distribution(oracle):
    d = np.stack([oracle.draws(4357, i, k) for i in range(4000) for k in (0, 1, 2, 3, 4)])
    u = d[:, [0, 1, 2, 3, 4]]
    assert (u >= 0).all() and (u < 1).all() and (d[:, 7] == 0).all()
    assert np.abs(u.mean(0) - 0.5).max() < 0.012 and np.abs(u.var(0) - 1 / 12).max() < 0.004
    # all seven values come from ONE Philox block: fields cut from the same word must not correlate
    cc = np.corrcoef(d[:, :7].T)
    assert np.abs(cc - np.eye(7)).max() < 0.03
    for g in (d[:, 5], d[:, 6]):
        assert abs(g.mean()) < 0.03 and abs(g.var() - 1) < 0.04 and abs((g ** 4).mean() - 3) < 0.3
    assert abs(np.corrcoef(d[:, 5], d[:, 6])[0, 1]) < 0.03
    # counter-based: same (seed, ray, k) -> same record; any change -> different record
    assert np.array_equal(oracle.draws(1, 2, 3), oracle.draws(1, 2, 3))
    assert not np.array_equal(oracle.draws(1, 2, 3), oracle.draws(1, 2, 4))
    assert not np.array_equal(oracle.draws(1, 2, 3), oracle.draws(2, 2, 3))
    assert not np.array_equal(oracle.draws(1, 2, 3), oracle.draws(1, 1 << 32 | 2, 3))


def test_f32_and_f64_modes_agree_per_ray(oracle):
    """The north-star's replay criterion between the two arithmetics of the oracle: same draws, per-ray escape
    port / status / bin equal except <= 1e-4 of the rays (those within FP32 epsilon of a boundary).  Rounding
    differences do not grow along a trajectory: the end-point error is independent of the chain length."""
    n = 300_000
    sc, src = oracle.scene(), oracle.source()
    a, sa = oracle.trace(sc, src, n, seed=21, prec=oracle.F32)
    b, sb = oracle.trace(sc, src, n, seed=21, prec=oracle.F64)
    same = (a["status"] == b["status"]) & (a["n_hits"] == b["n_hits"])
    assert 1 - same.mean() <= 1e-4, 1 - same.mean()
    assert (oracle.port_flags(sc, a) != oracle.port_flags(sc, b)).mean() <= 1e-4
    ex = same & (a["status"] == oracle.EXITED)
    dpos = np.abs(a["pos"][ex] - b["pos"][ex]).max(1)
    ddir = np.abs(a["dir"][ex] - b["dir"][ex]).max(1)
    assert dpos.max() < 0.05 and np.median(ddir) < 1e-6 and ddir.max() < 1e-3
    long_, short = b["n_hits"][ex] >= 150, b["n_hits"][ex] <= 10
    assert long_.sum() > 1000 and np.median(ddir[long_]) < 3 * np.median(ddir[short]) + 1e-7     # no growth
    m = oracle.map_spec(mode=oracle.MAP_DIRECTION)
    ca = oracle.map_records(sc, m, a).astype(np.int64)
    cb = oracle.map_records(sc, m, b).astype(np.int64)
    assert np.abs(ca - cb).sum() <= 2e-4 * cb.sum()                        # a moved ray changes two bins
    for key in ("n_exit_port", "n_absorbed", "n_bounces"):
        assert abs(sa[key] / sb[key] - 1) < 1e-4


def test_replay_of_own_tape_reproduces_trace(oracle):
    sc, src = oracle.scene(theta_max=164.0), oracle.source()
    n = 5000
    rec, _ = oracle.trace(sc, src, n, seed=5, prec=oracle.F32)
    tape, off = oracle.make_tape(sc, src, n, seed=5)
    assert (np.diff(off.astype(np.int64)) == rec["n_hits"]).all()
    ray0 = np.tile(np.array([-60.0, 0.0, -75.0, 5.0, 0.0, 0.0]), (n, 1))
    rep = oracle.replay(sc, ray0, tape, off, prec=oracle.F32)
    assert rep.tobytes() == rec.tobytes()
    # a truncated tape ends in TAPE_END, an empty one too
    off2 = off.copy(); off2[1:] = np.minimum(off2[1:], off2[:-1] + 1)
    off2 = np.concatenate([[0], np.cumsum(np.minimum(np.diff(off.astype(np.int64)), 1))]).astype(np.uint64)
    idx = off[:-1][np.diff(off.astype(np.int64)) > 0].astype(np.int64)
    rep2 = oracle.replay(sc, ray0, tape[idx], off2, prec=oracle.F32)
    assert ((rep2["status"] == oracle.TAPE_END) | (rep2["n_hits"] == 1)).all()


def test_map_modes_consistency(oracle):
    """LINE map: F32 division-free test vs the literal F64 formula differ only for pairs on the disk rim."""
    sc, src = oracle.scene(), oracle.source()
    rec, st = oracle.trace(sc, src, 20_000, seed=1, prec=oracle.F32)
    for mode in (oracle.MAP_LINE, oracle.MAP_TRACEONCE_COMPAT):
        a = oracle.map_records(sc, oracle.map_spec(mode=mode), rec, prec=oracle.F32).astype(np.int64)
        b = oracle.map_records(sc, oracle.map_spec(mode=mode), rec, prec=oracle.F64).astype(np.int64)
        assert a.sum() > 1_000_000
        assert np.abs(a - b).sum() <= 2e-5 * a.sum()
    # DIRECTION: one bin per escaping ray with dz < 0
    d = oracle.map_records(sc, oracle.map_spec(mode=oracle.MAP_DIRECTION), rec)
    flags = oracle.port_flags(sc, rec)
    assert d.sum() == (flags & (rec["dir"][:, 2] < 0)).sum() == st["n_exit_port"]
    # pole / equator / wrap-around bins
    m = oracle.map_spec(mode=oracle.MAP_DIRECTION)
    def b(v):
        v = np.array(v, dtype=np.float32)
        return oracle.lib().orc_direction_bin(C.byref(m), v.ctypes.data_as(C.POINTER(C.c_float)))
    assert b([0, 0, -1]) == 0 and b([0, 0, 1]) == -1 and b([1, 0, 0]) == -1
    assert b([1, 0, -1e-6]) == 179 * 90 and b([1, -1e-6, -1e-6]) == 179 * 90 + 89
    v = np.array([-1.0, 1e-3, -1.2]) / np.linalg.norm([-1.0, 1e-3, -1.2])
    th, ph = np.degrees(np.arccos(-v[2])), np.degrees(np.arctan2(v[1], v[0])) % 360
    assert b(v) == int(th / 0.5) * 90 + int(ph / 4.0) and b(v) % 90 == 44


def test_detector_pose_is_the_references_misoriented_normal(oracle):
    """Detector::setPosition (fluxAtObserverFast.C:61-80): normal = (-dy, dx, dz)/|d|, not the radial direction."""
    p, n = (C.c_double * 3)(), (C.c_double * 3)()
    oracle.lib().orc_detector_pose(30.0, 90.0, 100.0, p, n)
    assert np.allclose(list(p), [0.0, 50.0, -100 - 100 * np.cos(np.radians(30))], atol=1e-9)
    assert np.allclose(list(n), [-0.5, 0.0, -np.cos(np.radians(30))], atol=1e-9)
    # a vertical line through the detector centre hits, one 21 cm away in the plane misses (width 40)
    L, v = (C.c_double * 3)(0.0, 50.0, 0.0), (C.c_double * 3)(0.0, 0.0, -1.0)
    assert oracle.lib().orc_detector_hit(p, n, 40.0, L, v) == 1
    L2 = (C.c_double * 3)(0.0, 71.1, 0.0)
    assert oracle.lib().orc_detector_hit(p, n, 40.0, L2, v) == 0


def test_horizon_count_scales_with_roughness(oracle):
    """hits whose roughness-tilted normal no longer faces the incoming ray.  A tilt of g sigma passes the
    horizon when the incidence angle is within it of grazing; for Lambertian arrivals on a sphere P(cos < x) = x^2, so the
    fraction of hits grows like E[min(1, (g sigma)^2)] / 2 for small sigma: 4e-5 at the reference's 0.01 rad, ~8 % at the
    0.5 rad of fluxAtObserver.C:156."""
    src = oracle.source((-60, 0, -80), (5, 2, 0))
    frac = {}
    for sig in (0.0, 0.01, 0.1, 0.5):
        kw = dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=sig, max_bounces=10000)
        e, r, h = oracle.count_horizon(oracle.scene(**kw), src, 20_000, prec=oracle.F64)
        e32, r32, h32 = oracle.count_horizon(oracle.scene(**kw), src, 20_000, prec=oracle.F32)
        assert abs(e - e32) <= max(3, 1e-3 * e) and abs(h - h32) <= 1e-3 * max(h, 1)
        assert r <= e and r <= 20_000
        frac[sig] = e / h if h else 0.0
    assert frac[0.0] == 0.0
    assert 1e-5 < frac[0.01] < 1e-4 and 2e-3 < frac[0.1] < 6e-3 and 0.06 < frac[0.5] < 0.11, frac
    assert abs(frac[0.1] / frac[0.01] / 100.0 - 1) < 0.35          # ~ sigma^2
