"""The macro mirror end to end on the GPU: same entry points as the reference's ROOT macros, outputs checked
against the oracle and against the reference's file formats."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "altair-raytracing_b200")


@pytest.fixture(scope="module")
def M(altb, tmp_path_factory):
    C.CDLL(altb.library_path(), mode=C.RTLD_GLOBAL)
    L = C.CDLL(os.path.join(PKG, "libaltair_macros.so"))
    d = C.c_double
    for name in ("altbm_sweepDetector", "altbm_sweepDetectorTwofold", "altbm_sweepDetectorTraceOnce"):
        getattr(L, name).argtypes = [C.c_int, C.c_char_p, C.c_int, d, d, d, d, d, d, d]
    L.altbm_set.argtypes = [C.c_char_p, d]
    L.altbm_last_csv.restype = C.c_char_p
    L.altbm_last_count.restype = C.c_longlong
    L.altbm_last_hist.restype = d
    L.altbm_last_fluxmap.restype = d
    out = tmp_path_factory.mktemp("macros")
    L.altbm_set_output_dir(str(out).encode())
    L.altbm_set(b"verbose", 0)
    L.altbm_set(b"advance_ray_ids", 0)        # the oracle comparisons below trace ray ids 0 .. n-1; see test_repeated_calls_*
    L.altbm_last_count.argtypes = [C.c_char_p]
    L.out = str(out)
    return L


def _rows(path):
    return np.array([[float(x) for x in l.split(",")] for l in open(path) if l[0].isdigit()])


def test_sweepDetectorTraceOnce(M, oracle):
    M.altbm_set(b"traceonce_rays", 20000)
    for shipped, mode in ((1, oracle.MAP_TRACEONCE_COMPAT), (0, oracle.MAP_LINE)):
        M.altbm_set(b"traceonce_as_shipped", shipped)
        M.altbm_sweepDetectorTraceOnce(0, b"res", 1, -60.0, 0.0, -75.0, 5.0, 0.0, 0.0, 164.0)
        path = M.altbm_last_csv().decode()
        assert os.path.basename(path).startswith("fluxmap_traceonce_20000rays_180x90_src-60_0_-75")
        rows = _rows(path)
        counts, st = oracle.fluxmap(oracle.scene(theta_max=164.0), oracle.source(), 20000, oracle.map_spec(mode=mode), seed=4357,
                                    prec=oracle.F32)
        assert np.array_equal(np.rint(rows[:, 2] * 20000).astype(np.uint64), counts)
        text = open(path).read()
        assert f"# Total rays exiting port: {st['n_exit_port']} out of 20000" in text
        assert "# Exit port angle: 164 degrees" in text and "# Ray tracing time:" in text and "# Detector sweep time:" in text
        assert M.altbm_last_fluxmap(1, 1) == counts[0] / 20000
    # the second run did not overwrite the first (getUniqueFilename)
    assert path.endswith("_1.csv")
    M.altbm_set(b"traceonce_as_shipped", 1)
    M.altbm_set(b"traceonce_rays", 100000)


def test_repeated_calls_are_independent_samples(M, oracle):
    """The reference's gRandom keeps advancing: the five repeats per port angle of sweepSeries (fluxAtObserverFast.C:1641-1673)
    are independent samples.  With advance_ray_ids = 1 (the default) every macro call takes the next block of ray ids."""
    n = 20000
    M.altbm_set(b"traceonce_rays", n)
    M.altbm_set(b"advance_ray_ids", 1)
    M.altbm_set(b"next_ray_id", 0)
    try:
        maps = []
        for rep in range(3):
            M.altbm_sweepDetectorTraceOnce(0, b"rep", 1, -60.0, 0.0, -75.0, 5.0, 0.0, 0.0, 170.0)
            assert M.altbm_last_count(b"first_ray_id") == rep * n
            rows = _rows(M.altbm_last_csv().decode())
            counts, _ = oracle.fluxmap(oracle.scene(theta_max=170.0), oracle.source(), n, oracle.map_spec(mode=oracle.MAP_TRACEONCE_COMPAT),
                                       seed=4357, ray_id0=rep * n, prec=oracle.F32)
            assert np.array_equal(np.rint(rows[:, 2] * n).astype(np.uint64), counts)
            maps.append(rows[:, 2])
        assert M.altbm_last_count(b"next_ray_id") == 3 * n
        assert not np.array_equal(maps[0], maps[1]) and not np.array_equal(maps[1], maps[2])
    finally:
        M.altbm_set(b"advance_ray_ids", 0)
        M.altbm_set(b"next_ray_id", 0)
        M.altbm_set(b"traceonce_rays", 100000)


def test_sweepDetector_per_position_and_twofold(M, oracle):
    M.altbm_set(b"rays_per_position", 300)
    M.altbm_set(b"n_theta_bins", 36); M.altbm_set(b"n_phi_bins", 18)
    try:
        M.altbm_sweepDetector(0, b"pp", 1, -60.0, 0.0, -75.0, 5.0, 0.0, 0.0, 170.0)
        path = M.altbm_last_csv().decode()
        rows = _rows(path)
        mp = oracle.map_spec(36, 18, 100.0, 40.0, oracle.MAP_PER_POSITION, rays_per_position=300)
        counts, st = oracle.fluxmap(oracle.scene(), oracle.source(), 36 * 18 * 300, mp, seed=4357, prec=oracle.F32)
        assert counts.sum() > 0 and np.array_equal(np.rint(rows[:, 2] * 300).astype(np.uint64), counts)
        assert f"# Total ray hits: {counts.sum()} out of {36 * 18 * 300}" in open(path).read()
        assert M.altbm_last_count(b"totalHitRays") == counts.sum()
        M.altbm_sweepDetectorTwofold(0, b"pp", 1, -60.0, 0.0, -75.0, 5.0, 2.0, 0.0, 170.0)
        rows = _rows(M.altbm_last_csv().decode())
        mp = oracle.map_spec(36, 18, 100.0, 40.0, oracle.MAP_TWOFOLD, rays_per_position=300)
        counts, _ = oracle.fluxmap(oracle.scene(), oracle.source(direction=(5.0, 2.0, 0.0)), 36 * 9 * 300, mp, seed=4357, prec=oracle.F32)
        got = np.zeros(36 * 18, dtype=np.uint64)
        for th, ph, fr in rows:
            got[int(th / 2.5) * 18 + int(ph / 20.0)] = round(fr * 300)
        assert np.array_equal(got, counts) and counts.sum() > 0
    finally:
        M.altbm_set(b"rays_per_position", 50000); M.altbm_set(b"n_theta_bins", 180); M.altbm_set(b"n_phi_bins", 90)


def test_makeIntegratingSphereNRays_and_distribution(M, oracle):
    M.altbm_set(b"nrays_macro", 50000)                      # BASELINE.json config C1
    M.altbm_makeIntegratingSphereNRays()
    sc = oracle.scene(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.0, max_bounces=10000, count_all_status=1)
    src = oracle.source((-60, 0, -80), (5, 0, 0))
    _, st = oracle.trace(sc, src, 50000, seed=4357, prec=oracle.F32, want_records=False)
    assert M.altbm_last_count(b"fluxCount") == st["n_exit_port"] and st["n_exit_port"] > 49900
    assert M.altbm_last_count(b"n_bounces") == st["n_bounces"]
    M.altbm_set(b"distribution_rays", 30000)
    M.altbm_distributionSphereDetectorSweep()
    rec, _ = oracle.trace(sc, src, 30000, seed=4357, prec=oracle.F32)
    esc = rec["pos"][:, 2] < -100.0
    dz = rec["dir"][esc].astype(np.float64)
    dz = dz[:, 2] / np.linalg.norm(dz, axis=1)
    h = np.histogram(dz, bins=100, range=(-1, 1))[0]
    got = np.array([M.altbm_last_hist(b"hDirectionZ", b) for b in range(1, 101)])
    assert np.abs(got - h).sum() <= 2 and got.sum() == esc.sum()        # (bin-edge ties in double vs numpy)
    assert M.altbm_last_count(b"fluxCount") == esc.sum()
    # as written in the reference, theta = sign(dx)*acos(dz) lies outside (-90,90): everything overflows (:94,48-50)
    assert sum(M.altbm_last_hist(b"hAngularDist", b) for b in range(1, 181)) == 0


def test_integratingSphereDetectorSweep(M, oracle):
    M.altbm_set(b"sweep_rays", 40000); M.altbm_set(b"sweep_dtheta", 5.0)
    M.altbm_integratingSphereDetectorSweep()
    path = M.altbm_last_csv().decode()
    lines = open(path).read().splitlines()
    assert lines[0] == "Theta(deg)\tPhi(deg)\tHitFraction" and len(lines) == 1 + 19 * 2
    sc = oracle.scene(theta_max=170.0, r_outer=105.0, world_half=200.0, reflectance=1.0, roughness=0.0, max_bounces=10000, count_all_status=1)
    rec, _ = oracle.trace(sc, oracle.source((-60, 0, -80), (5, 0, 0)), 40000, seed=4357, prec=oracle.F32)
    poses = [oracle.sweep_pose(t, p) for t in np.arange(-45, 45.01, 5.0) for p in (0.0, 180.0)]
    hits = oracle.disk_hits(sc, rec, np.array([c for c, _ in poses]), np.array([m for _, m in poses]), 5.0, 0.1)
    got = np.array([float(l.split("\t")[2]) for l in lines[1:]])
    assert np.allclose(got, hits / 40000, atol=1e-9) and hits.sum() > 0
    M.altbm_set(b"sweep_rays", 100000); M.altbm_set(b"sweep_dtheta", 0.5)


def test_cli_runner(M, altb):
    exe = os.path.join(PKG, "altb_macro")
    out = subprocess.run([exe, "makeIntegratingSphereNRays.C"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "Flux of rays through the exit port: 1000" in out.stdout or "Flux of rays through the exit port: 999" in out.stdout
    out = subprocess.run([exe, "nonLambertianFlux.C", "--set", "nonlambertian_rays=200", "--set", "verbose=0", "--out", M.out],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    rows = _rows(os.path.join(M.out, "fluxmap_data.csv"))
    assert rows.shape == (900, 3) and rows[:, 2].sum() > 0 and open(os.path.join(M.out, "fluxmap_data.csv")).readline() == "theta,phi,fraction\n"


def test_nonLambertianFlux_posthoc_and_per_bounce(M, oracle):
    """nonLambertianFlux::sweepDetector: by default the committed macro literally (post-hoc re-scatter, brdf_kind 3); with
    nonlambertian_posthoc = 0 the BRDF at every bounce (brdf_kind 1).  Both against the oracle, bin for bin."""
    rpp = 150
    base = dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.5, max_bounces=10000, count_all_status=1,
                brdf_param=(0.3, 0.4, 0.6, 0.0))
    mp = oracle.map_spec(45, 20, 100.0, 10.0, oracle.MAP_PER_POSITION, rays_per_position=rpp)
    M.altbm_set(b"nonlambertian_rays", rpp)
    maps = {}
    try:
        for posthoc, kind in ((1, 3), (0, 1)):
            M.altbm_set(b"nonlambertian_posthoc", posthoc)
            M.altbm_nonLambertianFlux_sweepDetector()
            rows = _rows(os.path.join(M.out, "fluxmap_data.csv"))
            counts, _ = oracle.fluxmap(oracle.scene(brdf_kind=kind, **base), oracle.source((-60, 0, -80), (5, 0, 0)), 900 * rpp, mp,
                                       seed=4357, prec=oracle.F32)
            assert rows.shape == (900, 3) and counts.sum() > 0
            assert np.array_equal(np.rint(rows[:, 2] * rpp).astype(np.uint64), counts), kind
            maps[kind] = counts
        assert not np.array_equal(maps[1], maps[3])
    finally:
        M.altbm_set(b"nonlambertian_rays", 100000); M.altbm_set(b"nonlambertian_posthoc", 1)


def test_macro_contract_setting(M, oracle):
    """Settings::contract: the macros run under the exact contract by default (CSV = oracle bit for bit, the tests above);
    with contract = 2 (fast arithmetic + Philox4x32-7) the same call gives another sample of the same map."""
    n = 400_000
    M.altbm_set(b"traceonce_rays", n)
    try:
        M.altbm_sweepDetectorTraceOnce(0, b"c0", 1, -60.0, 0.0, -75.0, 5.0, 0.0, 0.0, 170.0)
        a = _rows(M.altbm_last_csv().decode())[:, 2]
        M.altbm_set(b"contract", 2)
        M.altbm_sweepDetectorTraceOnce(0, b"c2", 1, -60.0, 0.0, -75.0, 5.0, 0.0, 0.0, 170.0)
        b = _rows(M.altbm_last_csv().decode())[:, 2]
        assert M.altbm_set(b"contract", 9) == 0          # accepted by the setter; the next call reports the library's error
        M.altbm_sweepDetectorTraceOnce(0, b"c9", 1, -60.0, 0.0, -75.0, 5.0, 0.0, 0.0, 170.0)
    finally:
        M.altbm_set(b"contract", 0); M.altbm_set(b"traceonce_rays", 100000)
    assert not np.array_equal(a, b)
    assert abs(b.sum() / a.sum() - 1) < 0.01              # 3.2e6 hits each, ~270 correlated hits per escaping ray
    k1, k2 = np.rint(a * n), np.rint(b * n)
    use = k1 + k2 > 200
    z = (k1[use] - k2[use]) / np.sqrt(k1[use] + k2[use])
    assert use.sum() > 5000 and (z ** 2).mean() < 1.8       # bins share rays (~60 effective degrees of freedom): no tighter than that


def test_full_size_sweepDetector_against_reference_map(M, altb):
    """BASELINE-size statistical parity on the GPU: the production macro (fluxAtObserverOptimize.C sweepDetector:
    16 200 positions x 50 000 fresh rays = 8.1e8 rays; 12 524 s in the reference's own footer) against the reference's
    committed map for the same parameters: per-bin Poisson z, chi2/ndf ~ 1, total hits."""
    import json
    import time
    G = os.path.join(ROOT, "tests", "golden")
    z = np.load(os.path.join(G, "perposition_170_dir5_0_0.npz"))
    k_ref = z["hits"].astype(float)
    t0 = time.time()
    M.altbm_sweepDetector(0, b"full", -1, -60.0, 0.0, -75.0, 5.0, 0.0, 0.0, 170.0)
    wall = time.time() - t0
    path = M.altbm_last_csv().decode()
    rows = _rows(path)
    assert rows.shape == (16200, 3)
    k = np.rint(rows[:, 2] * 50000)
    n = 50000.0
    p = (k + k_ref) / (2 * n)
    ok = p * 2 * n > 30
    zz = (k - k_ref)[ok] / np.sqrt(2 * n * p[ok] * (1 - p[ok]))
    chi2 = (zz ** 2).mean()
    text = open(path).read()
    print(f"full-size sweepDetector: wall {wall:.2f} s, chi2/ndf {chi2:.3f}, max|z| {np.abs(zz).max():.2f}, "
          f"hits {int(k.sum())} vs reference {int(k_ref.sum())}")
    assert 0.8 < chi2 < 1.3, chi2
    assert np.abs(zz).max() < 6.0
    assert abs(k.sum() / k_ref.sum() - 1) < 0.012          # known systematic: -0.44 % (profiles/r02_residual.json)
    assert "# Number of rays per position: 50000" in text and f"out of {16200 * 50000}" in text
    assert wall < 120
    json.dump({"wall_s": wall, "chi2_ndf": chi2, "max_abs_z": float(np.abs(zz).max()), "hits": int(k.sum()), "hits_reference": int(k_ref.sum()),
               "reference_seconds": 12523.9}, open(os.path.join(ROOT, "gpurun_out", "full_size_sweepDetector.json"), "w"))


def test_python_macro_wrapper(altb, oracle, tmp_path):
    from altair_raytracing_b200 import macros
    macros.set_output_dir(tmp_path)
    macros.set("verbose", 0); macros.set("traceonce_rays", 5000)
    macros.sweepDetectorTraceOnce(False, "py", 1, -60, 0, -75, 5, 0, 0, 170.0)
    rows = _rows(macros.last_csv())
    counts, _ = oracle.fluxmap(oracle.scene(), oracle.source(), 5000, oracle.map_spec(mode=oracle.MAP_TRACEONCE_COMPAT), prec=oracle.F32)
    assert np.array_equal(np.rint(rows[:, 2] * 5000).astype(np.uint64), counts)
    assert macros.last_fluxmap(1, 1) == counts[0] / 5000
    macros.set("traceonce_rays", 100000)
    with pytest.raises(KeyError):
        macros.set("no_such_setting", 1)


def test_full_size_statistical_parity_other_goldens(M, ctx, altb):
    """More of the reference's committed outputs at full size: the 163-degree production map (8 660 529 hits in its footer),
    the trace-once maps as shipped (semantics B) at 1e7 rays, and the escape counts of all three port sizes at 1e8 rays."""
    import json
    G = os.path.join(ROOT, "tests", "golden")
    gold = json.load(open(os.path.join(G, "golden.json")))
    out = {}
    # --- per-position, theta_max = 163
    z = np.load(os.path.join(G, "perposition_163_dir5_0_0.npz"))
    k_ref = z["hits"].astype(float)
    M.altbm_sweepDetector(0, b"full163", -1, -60.0, 0.0, -75.0, 5.0, 0.0, 0.0, 163.0)
    k = np.rint(_rows(M.altbm_last_csv().decode())[:, 2] * 50000)
    n = 50000.0
    p = (k + k_ref) / (2 * n)
    ok = p * 2 * n > 30
    zz = (k - k_ref)[ok] / np.sqrt(2 * n * p[ok] * (1 - p[ok]))
    out["perposition_163"] = {"chi2_ndf": float((zz ** 2).mean()), "max_abs_z": float(np.abs(zz).max()), "hits": int(k.sum()),
                              "hits_reference": int(k_ref.sum())}
    assert 0.8 < out["perposition_163"]["chi2_ndf"] < 1.3 and np.abs(zz).max() < 6.0 and abs(k.sum() / k_ref.sum() - 1) < 0.012    # known: -0.92 %
    # --- trace-once maps as shipped + escape fractions
    for theta in (160, 164, 170):
        zt = np.load(os.path.join(G, f"traceonce_{theta}.npz"))
        kr, nr = zt["hits"].astype(float), float(zt["n_rays"])
        n_big = 10_000_000
        c, st = ctx.trace_fluxmap(altb.scene(theta_max=float(theta)), altb.source(), n_big, altb.map_spec(mode=altb.MAP_TRACEONCE_COMPAT), seed=1)
        ratio = (c.sum() / n_big) / (kr.sum() / nr)
        _, st8 = ctx.trace_fluxmap(altb.scene(theta_max=float(theta)), altb.source(), 100_000_000, altb.map_spec(mode=altb.MAP_DIRECTION), seed=2)
        esc = st8[0]["n_exit_port"] / 1e8
        ref = np.array(gold["escape_counts"][str(theta)], dtype=float)
        esc_ref, esc_sig = ref.mean() / 1e5, ref.std(ddof=1) / 1e5 / np.sqrt(ref.size)
        out[f"theta_{theta}"] = {"traceonce_sum_ratio": float(ratio), "escape_fraction": esc, "escape_fraction_reference": esc_ref,
                                 "reference_sigma": esc_sig, "z": (esc - esc_ref) / esc_sig}
        assert abs(ratio - 1) < 0.015
        assert abs(esc - esc_ref) < 3 * esc_sig              # 3 sigma of the mean of the reference's own runs (5 or 10 of 1e5 rays)
    print(json.dumps(out))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "statistical_parity.json"), "w"), indent=1)


def test_residual_table_against_reference_maps(ctx, altb):
    """Tracks the one KNOWN systematic difference between the parity model and the reference: with
    2e9-ray LINE maps the reference's two complete per-position maps sit up to +-1.7 % off the model per 10-degree band of
    detector theta, 9-24 sigma of the reference's own Poisson error -- far below its per-bin resolution (every per-bin test
    passes) but systematic.  Nothing in this environment can resolve it (ROBAST internals); the table is written every round
    (-> profiles/rNN_residual.json) and the test fails if the residual GROWS beyond its known envelope."""
    import json
    G = os.path.join(ROOT, "tests", "golden")
    n = 2_000_000_000
    out = {"rays": n, "bands_deg": [[10 * b, 10 * b + 10] for b in range(9)], "note": "reference / model - 1 per band; sigma = reference Poisson"}
    for name, th in (("perposition_170_dir5_0_0", 170.0), ("perposition_163_dir5_0_0", 163.0)):
        k_ref = np.load(os.path.join(G, name + ".npz"))["hits"].astype(float).reshape(180, 90)
        c, st = ctx.trace_fluxmap(altb.scene(theta_max=th), altb.source(), n, altb.map_spec(mode=altb.MAP_LINE), seed=11)
        exp = 50000.0 * c[0].reshape(180, 90).astype(float) / n
        ratios, sig = [], []
        for b in range(9):
            r, e = k_ref[20 * b:20 * b + 20].sum(), exp[20 * b:20 * b + 20].sum()
            ratios.append(r / e - 1); sig.append((r - e) / np.sqrt(e))
        tot = k_ref.sum() / exp.sum() - 1
        chi2 = float((((k_ref - exp) ** 2 / np.maximum(exp, 1e-9))[exp > 15]).mean())
        out[f"theta_max_{int(th)}"] = {"ratio_minus_1": [float(x) for x in ratios], "sigma": [float(x) for x in sig], "total_ratio_minus_1": float(tot),
                                      "chi2_ndf_per_bin": chi2, "escape_fraction": st[0]["n_exit_port"] / n}
        assert max(abs(x) for x in ratios[:7]) < 0.025, ratios          # known envelope: -1.3 % ... +1.7 %
        assert abs(tot) < 0.012 and 0.9 < chi2 < 1.15, (tot, chi2)
    print(json.dumps(out))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "residual.json"), "w"), indent=1)
