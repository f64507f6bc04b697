"""The oracle against a THIRD, independently written implementation of SURVEY.md appendix A (tests/independent_model.py:
generic nearest-hit search, Rodrigues rotations, numpy's RNG).  The kernels equal the oracle bit for bit, so an error of the
algorithm common to kernel and oracle would show up here -- and only here -- as a statistical disagreement."""
import numpy as np

import independent_model as IM


def _z(k1, n1, k2, n2):
    p = (k1 + k2) / (n1 + n2)
    return (k1 / n1 - k2 / n2) / np.sqrt(max(p * (1 - p), 1e-12) * (1 / n1 + 1 / n2))


def _compare(oracle, n, kw_o, kw_m, src, direction, seed):
    rec, st = oracle.trace(oracle.scene(**kw_o), oracle.source(src, direction), n, seed=seed, prec=oracle.F64)
    m = IM.trace(n, src=src, direction=direction, seed=seed, **kw_m)
    assert (m["status"] > 0).all()
    # status fractions and the port test of fluxAtObserverOptimize.C:309,323 (exited AND last z < -100)
    port_m = np.count_nonzero((m["status"] == IM.EXITED) & (m["pos"][:, 2] < -100.0))
    zs = {"port": _z(port_m, n, st["n_exit_port"], n),
          "exited": _z(np.count_nonzero(m["status"] == IM.EXITED), n, st["n_exited"], n),
          "suspended": _z(np.count_nonzero(m["status"] == IM.SUSPENDED), n, st["n_suspended"], n)}
    # surface hits per ray (both sample the same geometric-like law: compare means with the pooled spread)
    h_o, h_m = rec["n_hits"].astype(np.float64), m["n_hits"].astype(np.float64)
    zs["hits"] = (h_m.mean() - h_o.mean()) / np.sqrt(h_m.var() / n + h_o.var() / n)
    return rec, st, m, zs


def test_c2_scene_statistics(oracle):
    """The production scene (fluxAtObserverFast.C:33-41,192-230): rho 0.99, sigma 0.01, theta_max 170, box 300."""
    n = 300_000
    kw_o = dict(theta_max=170.0)
    kw_m = dict(theta_max=170.0, reflectance=0.99, roughness=0.01, world_half=300.0, max_bounces=50000)
    rec, st, m, zs = _compare(oracle, n, kw_o, kw_m, (-60.0, 0.0, -75.0), (5.0, 0.0, 0.0), seed=11)
    assert all(abs(v) < 4.0 for v in zs.values()), zs
    # the reference's own footers: 42 579 escapes per 1e5 rays at 170 deg (trace_once_test_*/*.csv:16221)
    port_m = np.count_nonzero((m["status"] == IM.EXITED) & (m["pos"][:, 2] < -100.0)) / n
    assert abs(port_m - 0.42579) < 4 * np.sqrt(0.42579 * 0.57421 * (1 / n + 1 / 5e5)), port_m
    # port-edge hits: rare (0.019 per ray at 170 deg) but they carry the 0.8 % thick-wall effect
    assert 0.015 < m["edge_hits"].mean() < 0.023, m["edge_hits"].mean()
    # exit directions of the escaping rays, 20 bins in dz
    pm = (m["status"] == IM.EXITED) & (m["pos"][:, 2] < -100.0)
    po = (rec["status"] == oracle.EXITED) & (rec["pos"][:, 2] < -100.0)
    hm, _ = np.histogram(m["dir"][pm, 2], bins=20, range=(-1, 0))
    ho, _ = np.histogram(rec["dir"][po, 2].astype(np.float64), bins=20, range=(-1, 0))
    ok = hm + ho > 50
    chi2 = (((hm - ho * hm.sum() / ho.sum()) ** 2) / (hm + ho * (hm.sum() / ho.sum()) ** 2))[ok].sum() / (ok.sum() - 1)
    assert chi2 < 2.2, chi2
    # and the azimuth of the exit points on the box floor is NOT uniform (first-bounce hot spot): both see the same asymmetry
    fm = np.mean(m["pos"][pm, 0] > 0); fo = np.mean(rec["pos"][po, 0] > 0)
    assert abs(fm - fo) < 4 * np.sqrt(0.25 * (1 / pm.sum() + 1 / po.sum())), (fm, fo)


def test_large_roughness_and_unit_reflectance(oracle):
    """fluxAtObserver.C:147-160: sigma = 0.5 rad (tilted normals past the horizon: the mirror rule of DESIGN.md matters for
    8 % of the hits), rho = 1, box 200, limit 10000, source (-60, 0, -80), direction (5, 2, 0)."""
    n = 60_000
    kw_o = dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.5, max_bounces=10000)
    kw_m = dict(theta_max=170.0, reflectance=1.0, roughness=0.5, world_half=200.0, max_bounces=10000)
    rec, st, m, zs = _compare(oracle, n, kw_o, kw_m, (-60.0, 0.0, -80.0), (5.0, 2.0, 0.0), seed=5)
    assert all(abs(v) < 4.0 for v in zs.values()), zs
    assert st["n_exited"] == n and (m["status"] == IM.EXITED).all()
    assert abs(m["dir"][:, 2].mean() - rec["dir"][:, 2].astype(np.float64).mean()) < 4 * np.sqrt(2 * 0.056 / n)


def test_big_port_thick_shell_and_bounce_limit(oracle):
    """Other corners: theta_max 160 with the 4.9 cm deep rim of integratingSphereDetectorSweep.C:119 (edge hits are 10 x more
    frequent), and a bounce limit that suspends most rays."""
    n = 150_000
    kw_o = dict(theta_max=160.0, r_outer=105.0, world_half=200.0, reflectance=0.97, roughness=0.0, max_bounces=10000)
    kw_m = dict(theta_max=160.0, r_outer=105.0, reflectance=0.97, roughness=0.0, world_half=200.0, max_bounces=10000)
    rec, st, m, zs = _compare(oracle, n, kw_o, kw_m, (-60.0, 0.0, -75.0), (5.0, 0.0, 0.0), seed=3)
    assert all(abs(v) < 4.0 for v in zs.values()), zs
    assert m["edge_hits"].mean() > 0.04          # (0.02 with the 0.9 cm rim at 170 deg)
    kw_o = dict(theta_max=170.0, reflectance=1.0, max_bounces=12)
    kw_m = dict(theta_max=170.0, reflectance=1.0, roughness=0.01, world_half=300.0, max_bounces=12)
    rec, st, m, zs = _compare(oracle, 100_000, kw_o, kw_m, (-60.0, 0.0, -75.0), (5.0, 0.0, 0.0), seed=4)
    assert all(abs(v) < 4.0 for v in zs.values()), zs
    assert st["n_suspended"] > 80_000


def test_detector_map_restatement(oracle):
    """The in-repo half of the path (Detector::setPosition + checkIntersection, fluxAtObserverFast.C:61-119): the numpy
    restatement on the ORACLE's records must give the oracle's F64 LINE map hit for hit (rim cases aside), and on the
    independent model's records the same map statistically."""
    n = 120_000
    kw = dict(theta_max=170.0)
    rec, st = oracle.trace(oracle.scene(**kw), oracle.source(), n, seed=21, prec=oracle.F64)
    counts = oracle.map_records(oracle.scene(**kw), oracle.map_spec(mode=oracle.MAP_LINE), rec, prec=oracle.F64).reshape(180, 90)
    rows = [0, 1, 45, 100, 150, 179]
    po = (rec["status"] == oracle.EXITED) & (rec["pos"][:, 2] < -100.0)
    mine = IM.line_map(rec["pos"][po].astype(np.float64), rec["dir"][po].astype(np.float64), rows=rows)
    assert np.abs(mine - counts[rows]).sum() <= 3 and mine.sum() > 20_000, (np.abs(mine - counts[rows]).sum(), mine.sum())
    m = IM.trace(n, theta_max=170.0, seed=22)
    pm = (m["status"] == IM.EXITED) & (m["pos"][:, 2] < -100.0)
    # row totals; their spread comes from the data (a ray hits many bins of a row -- all 90 in the row at the pole): 10 chunks
    pos, dr = m["pos"][pm], m["dir"][pm]
    chunks = np.array([IM.line_map(pos[c::10], dr[c::10], rows=rows).sum(axis=1) for c in range(10)], dtype=np.float64)
    theirs, var = chunks.sum(axis=0), 10.0 * chunks.var(axis=0, ddof=1)
    for a in range(len(rows)):
        k2 = float(counts[rows[a]].sum())
        assert abs(theirs[a] - k2) < 5.0 * np.sqrt(2.0 * var[a] + 9), (rows[a], theirs[a], k2, var[a])
