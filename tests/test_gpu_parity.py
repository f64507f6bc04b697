"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: bit-exact records / integer maps against the oracle's F32 arithmetic-contract mode;
statistical agreement with the F64 physics mode and with the reference's goldens.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 4357


def _scenes(mod):
    """(name, scene kwargs) shared by oracle and product (same field layout)."""
    return {
        "c2_lambert_rough": dict(theta_max=170.0),                                   # fluxAtObserverFast.C:33-41
        "c1_rho1_sigma0": dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.0,
                               max_bounces=10000, count_all_status=1),               # makeIntegratingSphereNRays.C:25-39
        "c3_custom_mirror": dict(theta_max=170.0, brdf_kind=1, brdf_param=(0.3, 0.4, 0.6, 0.0)),
        "specular": dict(theta_max=170.0, lambertian=0, roughness=0.05, max_bounces=2000),
        "big_port_160": dict(theta_max=160.0),
        "thick_shell": dict(theta_max=170.0, r_outer=105.0, world_half=200.0, reflectance=1.0, roughness=0.0,
                            max_bounces=10000),                                      # integratingSphereDetectorSweep.C:119
        "rough_half": dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.5, max_bounces=10000),
        # edge cases of the domain
        "suspended_limit_7": dict(theta_max=170.0, reflectance=1.0, max_bounces=7),          # every ray hits SetLimit
        "limit_1": dict(theta_max=170.0, max_bounces=1),
        "tiny_port_178": dict(theta_max=178.0, count_all_status=1),    # wall points below z=-100: absorbed rays pass the z test
        "huge_port_100": dict(theta_max=100.0, roughness=0.0),
        "black_wall": dict(theta_max=170.0, reflectance=0.0),          # every ray absorbed at its first hit
        "all_specular_lobe": dict(theta_max=170.0, brdf_kind=1, brdf_param=(1.0, 1.0, 0.0, 0.0)),
        "all_diffuse_lobe": dict(theta_max=170.0, brdf_kind=1, brdf_param=(0.3, 0.0, 1.0, 0.0), roughness=0.0),
        "mirror_sphere": dict(theta_max=170.0, lambertian=0, roughness=0.0, reflectance=0.999, max_bounces=3000),
        "small_world": dict(theta_max=165.0, world_half=102.0),
        "cos2_lobe": dict(theta_max=170.0, brdf_kind=2, brdf_param=(2.0, 60.0, 0.0, 0.0)),     # 'nonLambertianFlux copy.C':31-70
        "cos5_lobe_smooth": dict(theta_max=166.0, brdf_kind=2, brdf_param=(5.0, 45.0, 0.0, 0.0), roughness=0.0, reflectance=1.0,
                                 max_bounces=10000),
        # brdf_kind 3: the committed nonLambertianFlux.C literally (:213-226 scene, :246-268 re-scatter + second trace)
        "c1_posthoc_macro": dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.5, max_bounces=10000,
                                 count_all_status=1, brdf_kind=3, brdf_param=(0.3, 0.4, 0.6, 0.0)),
        # ... small world, thick shell, all-specular lobe: 11 % of the second rays meet the shell again (outer surface, port
        # edge, cavity: up to ~50 more hits), absorbed primaries stay as they are
        "posthoc_stress": dict(theta_max=140.0, world_half=103.0, r_outer=102.5, reflectance=0.95, roughness=0.2,
                               count_all_status=1, brdf_kind=3, brdf_param=(1.0, 1.0, 0.0, 0.0)),
        # ... the source looks straight out of the port (every primary is the same record), no roughness
        "posthoc_all_exit": dict(theta_max=120.0, world_half=110.0, reflectance=0.9, roughness=0.0, brdf_kind=3,
                                 brdf_param=(0.3, 0.4, 0.6, 0.0)),
    }


def _records_equal(a, b):
    return (np.array_equal(a["pos"].view(np.uint32), b["pos"].view(np.uint32))
            and np.array_equal(a["dir"].view(np.uint32), b["dir"].view(np.uint32))
            and np.array_equal(a["n_hits"], b["n_hits"]) and np.array_equal(a["status"], b["status"]))


def test_draws_bit_exact(ctx, oracle):
    for k in (0, 1, 77, 49999):
        g = ctx.draws(SEED, 1 << 33, 4096, k)
        o = np.stack([oracle.draws(SEED, (1 << 33) + i, k) for i in range(4096)])
        assert np.array_equal(g.view(np.uint32), o.view(np.uint32)), f"k={k}"


def test_f32_primitives_are_ieee_exact(ctx, oracle):
    """The written-out sqrt / reciprocal fast paths must be THE correctly rounded results (what sqrtf and 1.0f/x give on
    the CPU) on every argument the kernels can feed them: exhaustive over the 2^24 + 1 uniforms k 2^-24 (sqrt(u_r),
    sqrt(1 - u_r)), over every f32 in [1, 2] (orthonormal-basis reciprocal) and dense samples of the other ranges; the
    log and azimuth tables must equal the oracle's restatement bit for bit."""
    import ctypes as C
    rng = np.random.default_rng(3)
    u = (np.arange((1 << 24) + 1, dtype=np.float64) * 2.0 ** -24).astype(np.float32)
    assert np.array_equal(ctx.probe_f32(0, u).view(np.uint32), np.sqrt(u).view(np.uint32))
    x = np.concatenate([rng.uniform(0.25, 4.0, 4_000_000), 10.0 ** rng.uniform(-6.5, 1.5, 4_000_000)]).astype(np.float32)
    assert np.array_equal(ctx.probe_f32(0, x).view(np.uint32), np.sqrt(x).view(np.uint32))
    m = (np.arange((1 << 23) + 1, dtype=np.uint32) + np.uint32(0x3f800000)).view(np.float32)      # every f32 in [1, 2]
    for v in (m, -m, rng.uniform(0.4, 2.5, 4_000_000).astype(np.float32)):
        assert np.array_equal(ctx.probe_f32(1, v).view(np.uint32), (np.float32(1.0) / v).view(np.uint32))
    L = oracle.lib()
    t = np.concatenate([rng.integers(1, (1 << 20) + 1, 60000), [1, 2, 3, (1 << 19) - 1, 1 << 19, (1 << 20) - 1, 1 << 20]])
    want = np.array([L.orc_log_u20(int(v)) for v in t], dtype=np.float32)
    assert np.array_equal(ctx.probe_f32(2, t.astype(np.float32)).view(np.uint32), want.view(np.uint32))
    q = np.concatenate([rng.integers(0, 1 << 20, 60000), np.arange(0, 1 << 20, 128)[:2000], [0, 127, (1 << 20) - 1]]).astype(np.uint32)
    s_, c_ = C.c_float(), C.c_float()
    ws, wc = np.zeros(q.size, np.float32), np.zeros(q.size, np.float32)
    for i, v in enumerate(q):
        L.orc_sincos2pi_q20(int(v), C.byref(s_), C.byref(c_))
        ws[i], wc[i] = s_.value, c_.value
    assert np.array_equal(ctx.probe_f32(3, q.astype(np.float32)).view(np.uint32), ws.view(np.uint32))
    assert np.array_equal(ctx.probe_f32(4, q.astype(np.float32)).view(np.uint32), wc.view(np.uint32))


def test_lobe_draws_bit_exact(ctx, oracle):
    import ctypes as C
    g = ctx.draws(SEED, 77, 4096, 3, lobe_n=2, lobe_deg=60.0)
    o = np.zeros((4096, 8), dtype=np.float32)
    buf = (C.c_float * 8)()
    for i in range(4096):
        oracle.lib().orc_draws_lobe(SEED, 77 + i, 3, 2, np.float32(60.0 * np.pi / 180.0), buf)
        o[i] = list(buf)
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32))
    plain = ctx.draws(SEED, 77, 4096, 3)
    assert np.array_equal(plain[:, [0, 2, 3, 4, 5, 6]], g[:, [0, 2, 3, 4, 5, 6]]) and not np.array_equal(plain[:, 1], g[:, 1])


@pytest.mark.parametrize("name", ["c2_lambert_rough", "c1_rho1_sigma0", "c3_custom_mirror", "specular",
                                  "big_port_160", "thick_shell", "rough_half", "suspended_limit_7", "limit_1",
                                  "tiny_port_178", "huge_port_100", "black_wall", "all_specular_lobe",
                                  "all_diffuse_lobe", "mirror_sphere", "small_world", "cos2_lobe", "cos5_lobe_smooth",
                                  "c1_posthoc_macro", "posthoc_stress", "posthoc_all_exit"])
def test_trace_records_bit_exact(ctx, oracle, altb, name):
    kw = _scenes(altb)[name]
    n = 100_000
    src = (-60.0, 0.0, -75.0) if "c1" not in name else (-60.0, 0.0, -80.0)
    direction = (5.0, 0.0, 0.0) if "mirror" not in name else (5.0, 4.0, 1.0)       # (5,2,0),(5,4,0): fluxAtObserverOptimize.C:908-912
    g_rec, g_st = ctx.trace_records(altb.scene(**kw), altb.source(src, direction), n, seed=SEED)
    o_rec, o_st = oracle.trace(oracle.scene(**kw), oracle.source(src, direction), n, seed=SEED, prec=oracle.F32)
    if name == "suspended_limit_7":
        assert (g_rec["status"] != altb.ABSORBED).all() and (g_rec["status"] == altb.SUSPENDED).sum() > 0.8 * n
        assert g_rec["n_hits"].max() == 7
    if name == "black_wall":
        assert (g_rec["status"] == altb.ABSORBED).all() and (g_rec["n_hits"] == 1).all()
    if name == "tiny_port_178":
        assert ((g_rec["status"] == altb.ABSORBED) & (g_rec["pos"][:, 2] < -100.0)).sum() > 0
    bad = np.flatnonzero((g_rec["status"] != o_rec["status"]) | (g_rec["n_hits"] != o_rec["n_hits"])
                         | (g_rec["pos"].view(np.uint32) != o_rec["pos"].view(np.uint32)).any(axis=1)
                         | (g_rec["dir"].view(np.uint32) != o_rec["dir"].view(np.uint32)).any(axis=1))
    assert bad.size == 0, f"{bad.size} of {n} rays differ, first {bad[:5]}: {g_rec[bad[:2]]} vs {o_rec[bad[:2]]}"
    for key in ("n_rays", "n_exited", "n_exit_port", "n_absorbed", "n_suspended", "n_bounces"):
        assert g_st[key] == o_st[key], key
    assert g_st["n_exited"] + g_st["n_absorbed"] + g_st["n_suspended"] == n      # conservation
    if name == "posthoc_stress":     # the second trace really ran: more hits than the plain Lambertian trace of the same ids
        p_rec, _ = ctx.trace_records(altb.scene(**dict(kw, brdf_kind=0)), altb.source(src, direction), n, seed=SEED)
        extra = g_rec["n_hits"].astype(np.int64) - p_rec["n_hits"].astype(np.int64)
        assert (extra >= 0).all() and (extra == 1).sum() > 2000 and (extra > 1).sum() > 2000


@pytest.mark.parametrize("mode", ["LINE", "PER_POSITION", "DIRECTION"])
def test_posthoc_fluxmap_bit_exact(ctx, oracle, altb, mode):
    """brdf_kind 3 through the map stage: the 45x20 / 10 cm map of nonLambertianFlux.C:307-387 (fresh rays per position),
    the same as one LINE map, and a DIRECTION map (all through the record path), ragged batches."""
    kw = dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.5, max_bounces=10000, count_all_status=1,
              brdf_kind=3, brdf_param=(0.3, 0.4, 0.6, 0.0))
    rpp = 40
    n = 45 * 20 * rpp if mode == "PER_POSITION" else 30_011
    gm = altb.map_spec(45, 20, 100.0, 10.0, getattr(altb, "MAP_" + mode), rays_per_position=rpp)
    om = oracle.map_spec(45, 20, 100.0, 10.0, getattr(oracle, "MAP_" + mode), rays_per_position=rpp)
    ctx.set_batch(7_001)
    try:
        g_counts, g_st = ctx.trace_fluxmap(altb.scene(**kw), altb.source((-60, 0, -80)), n, gm, seed=9)
    finally:
        ctx.set_batch(0)
    o_counts, o_st = oracle.fluxmap(oracle.scene(**kw), oracle.source((-60, 0, -80)), n, om, seed=9, prec=oracle.F32)
    assert np.array_equal(g_counts[0], o_counts) and o_counts.sum() > 0
    for key in ("n_rays", "n_exited", "n_exit_port", "n_absorbed", "n_suspended", "n_bounces"):
        assert g_st[0][key] == o_st[key], key
    with pytest.raises(Exception):
        ctx.trace_paths(altb.scene(**kw), altb.source((-60, 0, -80)), 4, 16)


@pytest.mark.parametrize("kw", [dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.5, max_bounces=10000),   # fluxAtObserver.C:147-160
                                dict(theta_max=170.0), dict(theta_max=164.0, roughness=0.2, brdf_kind=1),
                                dict(theta_max=170.0, lambertian=0, roughness=0.3, max_bounces=500),
                                dict(theta_max=170.0, roughness=0.0)])
def test_horizon_count_bit_exact(ctx, oracle, altb, kw):
    """surface hits whose roughness-tilted normal no longer faces the incoming ray, counted separately."""
    n = 60_000
    got = ctx.count_horizon(altb.scene(**kw), altb.source((-60, 0, -80), (5, 2, 0)), n, seed=SEED, ray_id0=17)
    want = oracle.count_horizon(oracle.scene(**kw), oracle.source((-60, 0, -80), (5, 2, 0)), n, seed=SEED, ray_id0=17, prec=oracle.F32)
    assert got == want, (got, want)
    if kw.get("roughness", 0.01) >= 0.2:
        assert got[0] > 1000 and got[1] > 1000
    if kw.get("roughness", 0.01) == 0.0:
        assert got == (0, 0, 0)


def test_ray_id_offsets_compose(ctx, altb):
    sc, src = altb.scene(), altb.source()
    full, _ = ctx.trace_records(sc, src, 50_000, seed=SEED, ray_id0=123)
    a, _ = ctx.trace_records(sc, src, 20_000, seed=SEED, ray_id0=123)
    b, _ = ctx.trace_records(sc, src, 30_000, seed=SEED, ray_id0=123 + 20_000)
    assert _records_equal(full, np.concatenate([a, b]))


@pytest.mark.parametrize("mode,n", [("DIRECTION", 300_000), ("LINE", 30_000), ("TRACEONCE_COMPAT", 30_000)])
@pytest.mark.parametrize("name", ["c2_lambert_rough", "c3_custom_mirror"])
def test_fluxmap_bit_exact(ctx, oracle, altb, name, mode, n):
    kw = _scenes(altb)[name]
    gm = altb.map_spec(mode=getattr(altb, "MAP_" + mode))
    om = oracle.map_spec(mode=getattr(oracle, "MAP_" + mode))
    g_counts, g_st = ctx.trace_fluxmap(altb.scene(**kw), altb.source(), n, gm, seed=SEED)
    o_counts, o_st = oracle.fluxmap(oracle.scene(**kw), oracle.source(), n, om, seed=SEED, prec=oracle.F32)
    assert g_counts.shape == (1, 16200)
    diff = np.flatnonzero(g_counts[0] != o_counts)
    assert diff.size == 0, f"{diff.size} bins differ; sum gpu {g_counts.sum()} oracle {o_counts.sum()}"
    for key in ("n_rays", "n_exited", "n_exit_port", "n_absorbed", "n_suspended", "n_bounces"):
        assert g_st[0][key] == o_st[key], key
    assert g_counts.sum() > 0


def test_fluxmap_small_odd_grid_and_batches(ctx, oracle, altb):
    """45x20 / 10 cm detector of nonLambertianFlux.C:320-327, ragged batch sizes."""
    kw = dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.5, max_bounces=10000)
    gm = altb.map_spec(45, 20, 100.0, 10.0, altb.MAP_LINE)
    om = oracle.map_spec(45, 20, 100.0, 10.0, oracle.MAP_LINE)
    n = 20_011
    ctx.set_batch(7_001)
    try:
        g_counts, g_st = ctx.trace_fluxmap(altb.scene(**kw), altb.source((-60, 0, -80)), n, gm, seed=9)
    finally:
        ctx.set_batch(0)
    o_counts, o_st = oracle.fluxmap(oracle.scene(**kw), oracle.source((-60, 0, -80)), n, om, seed=9, prec=oracle.F32)
    assert np.array_equal(g_counts[0], o_counts)
    assert g_st[0]["n_bounces"] == o_st["n_bounces"]


@pytest.mark.parametrize("mode", ["PER_POSITION", "TWOFOLD"])
def test_per_position_modes_bit_exact(ctx, oracle, altb, mode):
    """Fresh rays per detector position (fluxAtObserverOptimize.C:542-579) / shared by a 180-deg pair (Fast.C:660-720)."""
    rpp, nt, npb = 700, 30, 12
    groups = nt * npb if mode == "PER_POSITION" else nt * npb // 2
    n = groups * rpp + 123                                   # trailing rays beyond the last group are ignored
    gm = altb.map_spec(nt, npb, 100.0, 40.0, getattr(altb, "MAP_" + mode), rays_per_position=rpp)
    om = oracle.map_spec(nt, npb, 100.0, 40.0, getattr(oracle, "MAP_" + mode), rays_per_position=rpp)
    ctx.set_batch(50_001)                                    # groups straddle batch boundaries
    try:
        g_counts, g_st = ctx.trace_fluxmap(altb.scene(), altb.source(), n, gm, seed=SEED)
    finally:
        ctx.set_batch(0)
    o_counts, o_st = oracle.fluxmap(oracle.scene(), oracle.source(), n, om, seed=SEED, prec=oracle.F32)
    assert np.array_equal(g_counts[0], o_counts) and o_counts.sum() > 100
    assert g_st[0]["n_bounces"] == o_st["n_bounces"]


def test_multi_scene_and_empty(ctx, oracle, altb):
    thetas = (160.0, 164.0, 175.0)
    gm = altb.map_spec(mode=altb.MAP_DIRECTION)
    g_counts, g_st = ctx.trace_fluxmap([altb.scene(theta_max=t) for t in thetas], altb.source(), 40_000, gm, seed=SEED)
    for i, t in enumerate(thetas):
        o_counts, o_st = oracle.fluxmap(oracle.scene(theta_max=t), oracle.source(), 40_000,
                                        oracle.map_spec(mode=oracle.MAP_DIRECTION), seed=SEED, prec=oracle.F32)
        assert np.array_equal(g_counts[i], o_counts), t
        assert g_st[i]["n_exit_port"] == o_st["n_exit_port"]
    z_counts, z_st = ctx.trace_fluxmap(altb.scene(), altb.source(), 0, gm)
    assert z_counts.sum() == 0 and z_st[0]["n_rays"] == 0


def _direction_oracle(oracle, kw, n, seed=SEED, ray_id0=0, src=None):
    return oracle.fluxmap(oracle.scene(**kw), src or oracle.source(), n, oracle.map_spec(mode=oracle.MAP_DIRECTION), seed=seed,
                          ray_id0=ray_id0, prec=oracle.F32)


def test_batched_scenes_one_launch(ctx, oracle, altb):
    """Port-angle series (fluxAtObserverFast.C:1641-1673): scenes that differ only in theta_max share ONE persistent launch,
    the scene index is part of the claimed work unit; maps and statistics must equal the oracle's scene by scene."""
    thetas = [100.0, 137.5, 139.0, 150.0, 160.0, 163.0, 164.0, 166.0, 169.0, 170.0, 172.0, 175.0, 178.0, 179.5]
    n = 30_011
    gm = altb.map_spec(mode=altb.MAP_DIRECTION)
    l0, t0 = ctx.launches, ctx.trace_launches
    g_counts, g_st = ctx.trace_fluxmap([altb.scene(theta_max=t) for t in thetas], altb.source(), n, gm, seed=SEED, ray_id0=77)
    walls = sum(1 for t in thetas if t >= 139.0)                 # below 138.5 deg the pencil beam leaves through the port at once
    assert ctx.trace_launches - t0 == 1, "the scenes whose first event is the wall must share one k_trace launch"
    for i, t in enumerate(thetas):
        o_counts, o_st = _direction_oracle(oracle, dict(theta_max=t), n, ray_id0=77)
        assert np.array_equal(g_counts[i], o_counts), t
        for key in ("n_rays", "n_exited", "n_exit_port", "n_absorbed", "n_suspended", "n_bounces"):
            assert g_st[i][key] == o_st[key], (t, key)
    # theta_max = 100: the beam leaves untouched, through the SIDE of the world box (z = -75 > exit_z): exited, not "port"
    assert walls == 12 and g_st[0]["n_bounces"] == 0 and g_st[0]["n_exited"] == n and g_st[0]["n_exit_port"] == 0
    # the same through per-scene launches (ALTB_NO_BATCH is read per call)
    import os
    os.environ["ALTB_NO_BATCH"] = "1"
    try:
        t0 = ctx.trace_launches
        s_counts, s_st = ctx.trace_fluxmap([altb.scene(theta_max=t) for t in thetas], altb.source(), n, gm, seed=SEED, ray_id0=77)
        assert ctx.trace_launches - t0 == walls
    finally:
        del os.environ["ALTB_NO_BATCH"]
    assert np.array_equal(s_counts, g_counts) and [a["n_bounces"] for a in s_st] == [a["n_bounces"] for a in g_st]


def test_batched_scenes_160_and_mixed_groups(ctx, oracle, altb):
    """BASELINE config C5 shape: 160 port angles in one call; plus a call that mixes scenes which cannot share a launch
    (different reflectance / BRDF / count_all) with ones that can -- each group goes its own way, results per scene."""
    thetas = [100.0 + 0.5 * k for k in range(160)]
    n = 4_003
    gm = altb.map_spec(mode=altb.MAP_DIRECTION)
    t0 = ctx.trace_launches
    g_counts, g_st = ctx.trace_fluxmap([altb.scene(theta_max=t) for t in thetas], altb.source(), n, gm, seed=11)
    # 100 .. 137.5 deg: the beam leaves at once (no launch); 138.0 and 138.5 deg: its first event is the port rim (generic
    # one-thread-per-ray tracer, one launch each); the other 81 scenes share one persistent launch
    assert ctx.trace_launches - t0 == 3
    for i in (0, 75, 76, 77, 78, 100, 140, 159):
        o_counts, o_st = _direction_oracle(oracle, dict(theta_max=thetas[i]), n, seed=11)
        assert np.array_equal(g_counts[i], o_counts), thetas[i]
        assert g_st[i]["n_bounces"] == o_st["n_bounces"] and g_st[i]["n_absorbed"] == o_st["n_absorbed"]
    tot = sum(s["n_exited"] + s["n_absorbed"] + s["n_suspended"] for s in g_st)
    assert tot == 160 * n and all(s["n_rays"] == n for s in g_st)
    kws = [dict(theta_max=165.0), dict(theta_max=170.0, reflectance=0.95), dict(theta_max=172.0),
           dict(theta_max=170.0, brdf_kind=1), dict(theta_max=178.0, count_all_status=1), dict(theta_max=168.0, brdf_kind=1),
           dict(theta_max=171.0, roughness=0.0), dict(theta_max=150.0)]
    n = 20_000
    t0 = ctx.trace_launches
    g_counts, g_st = ctx.trace_fluxmap([altb.scene(**kw) for kw in kws], altb.source(), n, gm, seed=SEED)
    assert ctx.trace_launches - t0 == 5          # {165,172,150}, {0.95}, {CustomMirror 170,168}, {count_all: record path}, {sigma 0}
    for i, kw in enumerate(kws):
        o_counts, o_st = _direction_oracle(oracle, kw, n)
        assert np.array_equal(g_counts[i], o_counts), kw
        for key in ("n_rays", "n_exited", "n_exit_port", "n_absorbed", "n_suspended", "n_bounces"):
            assert g_st[i][key] == o_st[key], (kw, key)


def test_direction_sink_batches_and_id_windows(ctx, oracle, altb):
    """In-kernel binning with ragged launch sizes, and ray ids that straddle a multiple of 2^32 (the launch is split there:
    the high counter word is uniform inside a launch)."""
    gm = altb.map_spec(mode=altb.MAP_DIRECTION)
    kw = dict(theta_max=170.0, brdf_kind=1)
    n = 50_021
    ctx.set_batch(7_001)
    try:
        a_counts, a_st = ctx.trace_fluxmap([altb.scene(**kw), altb.scene(**dict(kw, theta_max=160.0))], altb.source(), n, gm, seed=5)
    finally:
        ctx.set_batch(0)
    for i, t in enumerate((170.0, 160.0)):
        o_counts, o_st = _direction_oracle(oracle, dict(kw, theta_max=t), n, seed=5)
        assert np.array_equal(a_counts[i], o_counts) and a_st[i]["n_bounces"] == o_st["n_bounces"]
    id0 = (1 << 32) - 20_000
    t0 = ctx.trace_launches
    g_counts, g_st = ctx.trace_fluxmap(altb.scene(**kw), altb.source(), n, gm, seed=5, ray_id0=id0)
    assert ctx.trace_launches - t0 == 2
    o_counts, o_st = _direction_oracle(oracle, kw, n, seed=5, ray_id0=id0)
    assert np.array_equal(g_counts[0], o_counts) and g_st[0]["n_bounces"] == o_st["n_bounces"]
    g_rec, _ = ctx.trace_records(altb.scene(**kw), altb.source(), n, seed=5, ray_id0=id0)
    o_rec, _ = oracle.trace(oracle.scene(**kw), oracle.source(), n, seed=5, ray_id0=id0, prec=oracle.F32)
    assert _records_equal(g_rec, o_rec)


def test_source_through_port_and_errors(ctx, oracle, altb):
    # a source aimed straight at the port: every ray leaves untouched
    src_g, src_o = altb.source((0, 0, -50), (0, 0, -1)), oracle.source((0, 0, -50), (0, 0, -1))
    g_rec, g_st = ctx.trace_records(altb.scene(), src_g, 1000)
    o_rec, _ = oracle.trace(oracle.scene(), src_o, 1000, prec=oracle.F32)
    assert _records_equal(g_rec, o_rec) and g_st["n_exit_port"] == 1000
    with pytest.raises(altb.AltbError):
        ctx.trace_records(altb.scene(), altb.source((0, 0, 150.0)), 10)          # outside the sphere
    with pytest.raises(altb.AltbError):
        ctx.trace_records(altb.scene(theta_max=45.0), altb.source(), 10)         # not a port
    with pytest.raises(altb.AltbError):
        ctx.trace_records(altb.scene(brdf_kind=7), altb.source(), 10)


def test_source_aimed_at_port_rim(ctx, oracle, altb):
    """First event = the conical port edge (between r_inner and r_outer): fresh rays do not start on the inner sphere,
    the library takes its generic one-thread-per-ray tracer; maps and records must still equal the oracle bit for bit."""
    kw = dict(theta_max=170.0, r_outer=105.0, world_half=200.0, reflectance=0.98, roughness=0.01, max_bounces=10000)
    th = np.deg2rad(170.0)
    q = 102.5 * np.array([np.sin(th), 0.0, np.cos(th)])            # a point on the edge, mid-wall
    p0 = np.array([0.0, 0.0, -50.0])
    for extra in (dict(), dict(brdf_kind=1, brdf_param=(0.3, 0.4, 0.6, 0.0))):
        k2 = dict(kw, **extra)
        g_src, o_src = altb.source(tuple(p0), tuple(q - p0)), oracle.source(tuple(p0), tuple(q - p0))
        g_rec, g_st = ctx.trace_records(altb.scene(**k2), g_src, 20_000, seed=SEED)
        o_rec, o_st = oracle.trace(oracle.scene(**k2), o_src, 20_000, seed=SEED, prec=oracle.F32)
        assert _records_equal(g_rec, o_rec)
        assert g_st["n_bounces"] == o_st["n_bounces"] and g_st["n_bounces"] > 20_000      # they do bounce (edge first)
        g_counts, _ = ctx.trace_fluxmap(altb.scene(**k2), g_src, 20_000, altb.map_spec(mode=altb.MAP_DIRECTION), seed=SEED)
        o_counts, _ = oracle.fluxmap(oracle.scene(**k2), o_src, 20_000, oracle.map_spec(mode=oracle.MAP_DIRECTION), seed=SEED,
                                     prec=oracle.F32)
        assert np.array_equal(g_counts[0], o_counts)
    with pytest.raises(altb.AltbError):
        ctx.trace_records(altb.scene(reflectance=-0.1), altb.source(), 10)


def test_replay_bit_exact(ctx, oracle, altb):
    kw = dict(theta_max=170.0)
    n = 20_000
    tape, off = oracle.make_tape(oracle.scene(**kw), oracle.source(), n, seed=SEED)
    ray0 = np.tile(np.array([-60.0, 0.0, -75.0, 5.0, 0.0, 0.0]), (n, 1))
    o_rec = oracle.replay(oracle.scene(**kw), ray0, tape, off, prec=oracle.F32)
    gm = altb.map_spec(mode=altb.MAP_DIRECTION)
    g_rec, g_bin, g_port = ctx.replay(altb.scene(**kw), ray0, tape, off, gm)
    assert _records_equal(g_rec, o_rec)
    # the replayed run equals the Philox run it was recorded from
    t_rec, _ = ctx.trace_records(altb.scene(**kw), altb.source(), n, seed=SEED)
    assert _records_equal(g_rec, t_rec)
    assert np.array_equal(g_port.astype(bool), oracle.port_flags(oracle.scene(**kw), o_rec))
    import ctypes as C
    om = oracle.map_spec(mode=oracle.MAP_DIRECTION)
    o_bin = np.array([oracle.lib().orc_direction_bin(C.byref(om), r["dir"].ctypes.data_as(C.POINTER(C.c_float)))
                      if p else -1 for r, p in zip(o_rec, g_port)], dtype=np.int32)
    assert np.array_equal(g_bin, o_bin)
    # ragged / perturbed starts and a truncated tape
    rng = np.random.default_rng(1)
    ray0b = ray0.copy()
    v = rng.normal(size=(n, 3))
    ray0b[:, :3] = v / np.linalg.norm(v, axis=1, keepdims=True) * (95.0 * rng.uniform(0, 1, (n, 1)) ** (1 / 3))
    ray0b[:, 3:] = rng.normal(size=(n, 3))                       # anywhere inside the cavity, any direction
    off2 = off.copy()
    cut = (off2[1:] - off2[:-1]) > 3
    lens = (off2[1:] - off2[:-1]).astype(np.int64)
    lens[cut & (np.arange(n) % 7 == 0)] = 3
    # re-pack offsets (tape rows stay where they were: use per-ray slices through a gather)
    new_off = np.zeros(n + 1, dtype=np.uint64); new_off[1:] = np.cumsum(lens)
    idx = np.concatenate([np.arange(int(off[i]), int(off[i]) + int(lens[i])) for i in range(n)])
    tape2 = tape[idx]
    o2 = oracle.replay(oracle.scene(**kw), ray0b, tape2, new_off, prec=oracle.F32)
    g2, _, _ = ctx.replay(altb.scene(**kw), ray0b, tape2, new_off, None)
    assert _records_equal(g2, o2)
    assert (g2["status"] == altb.TAPE_END).sum() > 0


def test_replay_against_double_precision_oracle(ctx, oracle, altb):
    """North-star replay criterion: feed recorded initial rays + draws, compare the FP32 GPU result with the
    DOUBLE-PRECISION oracle per ray: escape port flag, status and bin index exact, except <= 1e-4 of the rays."""
    n = 200_000
    for kw in (dict(theta_max=170.0), dict(theta_max=164.0, brdf_kind=1)):
        tape, off = oracle.make_tape(oracle.scene(**kw), oracle.source(), n, seed=2)
        ray0 = np.tile(np.array([-60.0, 0.0, -75.0, 5.0, 0.0, 0.0]), (n, 1))
        ref = oracle.replay(oracle.scene(**kw), ray0, tape, off, prec=oracle.F64)
        g_rec, g_bin, g_port = ctx.replay(altb.scene(**kw), ray0, tape, off, altb.map_spec(mode=altb.MAP_DIRECTION))
        ref_port = oracle.port_flags(oracle.scene(**kw), ref)
        import ctypes as C
        om = oracle.map_spec(mode=oracle.MAP_DIRECTION)
        ref_bin = np.array([oracle.lib().orc_direction_bin(C.byref(om), r["dir"].ctypes.data_as(C.POINTER(C.c_float)))
                            if p else -1 for r, p in zip(ref, ref_port)], dtype=np.int32)
        bad = (g_rec["status"] != ref["status"]) | (g_port.astype(bool) != ref_port) | (g_bin != ref_bin)
        assert bad.mean() <= 1e-4, (kw, bad.mean())
        assert (g_rec["status"] == altb.TAPE_END).sum() <= bad.sum()


def test_detector_sweep_bit_exact(ctx, oracle, altb):
    kw = dict(theta_max=170.0, r_outer=105.0, world_half=200.0, reflectance=1.0, roughness=0.0, max_bounces=10000)
    poses = [oracle.sweep_pose(t, p) for t in np.arange(-45, 45.01, 2.5) for p in (0.0, 180.0)]
    centers = np.array([c for c, _ in poses]); rots = np.array([m for _, m in poses])
    n = 50_000
    src = ((-60, 0, -80), (5, 0, 0))
    hits, st = ctx.detector_sweep(altb.scene(**kw), altb.source(*src), n, centers, rots, 5.0, 0.1, seed=SEED)
    o_rec, o_st = oracle.trace(oracle.scene(**kw), oracle.source(*src), n, seed=SEED, prec=oracle.F32)
    o_hits = oracle.disk_hits(oracle.scene(**kw), o_rec, centers, rots, 5.0, 0.1)
    assert np.array_equal(hits, o_hits)
    assert hits.sum() > 0 and st["n_bounces"] == o_st["n_bounces"]


def test_exit_rays_api(ctx, oracle, altb):
    kw = dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.0, max_bounces=10000, count_all_status=1)
    n = 20_000
    pos, d, npts, status, st = ctx.trace_exit_rays(altb.scene(**kw), altb.source((-60, 0, -80)), n, seed=SEED)
    o_rec, _ = oracle.trace(oracle.scene(**kw), oracle.source((-60, 0, -80)), n, seed=SEED, prec=oracle.F32)
    assert np.array_equal(pos.astype(np.float32), o_rec["pos"]) and np.array_equal(d.astype(np.float32), o_rec["dir"])
    assert np.array_equal(status, o_rec["status"].astype(np.uint8))
    assert np.array_equal(npts, 1 + o_rec["n_hits"] + (o_rec["status"] == 1))
    # rho = 1: every ray escapes (makeIntegratingSphereNRays.C prints fluxCount ~ n), mean exit dz -> -2/3
    assert st["n_exited"] == n
    dz = d[pos[:, 2] < -100.0][:, 2]
    assert abs(dz.mean() + 2.0 / 3.0) < 0.01


def test_statistics_vs_f64_physics_and_reference(ctx, oracle, altb):
    """Port flux fraction within 0.1 % of the F64 oracle; chi2/ndf ~ 1 on the direction map."""
    n = 4_000_000
    gm = altb.map_spec(mode=altb.MAP_DIRECTION)
    g_counts, g_st = ctx.trace_fluxmap(altb.scene(), altb.source(), n, gm, seed=11)
    o_counts, o_st = oracle.fluxmap(oracle.scene(), oracle.source(), n, oracle.map_spec(mode=oracle.MAP_DIRECTION),
                                    seed=12, prec=oracle.F64)        # independent seed, double precision
    fg, fo = g_st[0]["n_exit_port"] / n, o_st["n_exit_port"] / n
    sig = np.sqrt(2 * fo * (1 - fo) / n)
    assert abs(fg - fo) < 4 * sig and abs(fg / fo - 1) < 2e-3
    # reference golden: 42 579 +- 230 of 100 000 (trace_once_test_04_2-60_0_-75_5/*.csv:16221)
    assert abs(fg - 0.42579) < 4 * np.sqrt(0.42579 * 0.57421 / 5e5)
    # coarse 18x9 re-binning so that every cell has counts >> 1
    a = g_counts[0].reshape(18, 10, 9, 10).sum(axis=(1, 3)).astype(float)
    b = o_counts.reshape(18, 10, 9, 10).sum(axis=(1, 3)).astype(float)
    m = (a + b) > 50
    chi2 = (((a - b) ** 2) / (a + b))[m].sum() / m.sum()
    assert 0.6 < chi2 < 1.5, chi2


@pytest.mark.parametrize("name,kw,n", [
    ("c3", dict(theta_max=170.0, brdf_kind=1, brdf_param=(0.3, 0.4, 0.6, 0.0)), 16_000_000),
    # deep, narrow port channel: ~1/3 of the crossings hit the conical edge -> the out-of-line edge bounces and the resume queue work hard
    ("deep_port", dict(theta_max=176.0, r_outer=112.0, world_half=200.0, reflectance=0.995, roughness=0.01, max_bounces=10000), 4_000_000),
])
def test_large_run_bit_exact(ctx, oracle, altb, name, kw, n):
    """Tens of millions of rays (1.6e9 surface hits in total) against the oracle's F32 mode: flux map and every counter equal exactly, so the rare
    paths of the persistent kernel (queues, edge bounces, regeneration across chunks, kernel tail) are exercised at scale."""
    gm, om = altb.map_spec(mode=altb.MAP_DIRECTION), oracle.map_spec(mode=oracle.MAP_DIRECTION)
    g_counts, g_st = ctx.trace_fluxmap(altb.scene(**kw), altb.source(), n, gm, seed=SEED, ray_id0=1 << 34)
    o_counts, o_st = oracle.fluxmap(oracle.scene(**kw), oracle.source(), n, om, seed=SEED, ray_id0=1 << 34, prec=oracle.F32)
    assert np.array_equal(g_counts[0], o_counts)
    for key in ("n_rays", "n_exited", "n_exit_port", "n_absorbed", "n_suspended", "n_bounces"):
        assert g_st[0][key] == o_st[key], key


def test_full_size_properties(ctx, altb):
    """BASELINE-size launch (1e8 rays, C5 per-scene size): conservation and GPU-count invariance."""
    n = 100_000_000
    gm = altb.map_spec(mode=altb.MAP_DIRECTION)
    sc = altb.scene(brdf_kind=1)
    c_all, st = ctx.trace_fluxmap(sc, altb.source(), n, gm, seed=SEED)
    s = st[0]
    assert s["n_rays"] == n and s["n_exited"] + s["n_absorbed"] + s["n_suspended"] == n
    assert c_all.sum() <= s["n_exit_port"] and c_all.sum() > 0.999 * s["n_exit_port"]
    # sharded as 3 "ranks" with global ray ids: integer maps add up exactly
    parts = np.zeros_like(c_all)
    tot = 0
    for lo, hi in ((0, 33_333_333), (33_333_333, 70_000_000), (70_000_000, n)):
        c, stp = ctx.trace_fluxmap(sc, altb.source(), hi - lo, gm, seed=SEED, ray_id0=lo)
        parts += c
        tot += stp[0]["n_bounces"]
    assert np.array_equal(parts, c_all) and tot == s["n_bounces"]


def test_in_process_multi_device_context(altb, ctx):
    """altb_create(devices, n): one host thread drives several GPUs (what the C++ macros use: `threads` of the reference's
    sweepDetector = SetMaxThreads).  Rays are split by global id; the context owns the NCCL communicators and merges the
    per-device maps with one all-reduce (host sum with ALTB_NO_NCCL=1) -> identical to the single-device result, for the
    record path (LINE) and for the in-kernel direction sink with batched scenes."""
    import os
    import torch
    assert ctx.collective == "none"
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs in one process (covered by bench.py's rank-0 in-process probe at N > 1 and by the gloo test on CPU)")
    n = 3_000_001
    devs = list(range(torch.cuda.device_count()))
    scenes = [altb.scene(theta_max=t, brdf_kind=1) for t in (160.0, 170.0)]
    for mode in (altb.MAP_LINE, altb.MAP_DIRECTION):
        gm = altb.map_spec(mode=mode)
        one, st1 = ctx.trace_fluxmap(scenes, altb.source(), n, gm, seed=SEED)
        for no_nccl in (False, True):
            if no_nccl:
                os.environ["ALTB_NO_NCCL"] = "1"
            try:
                with altb.Context(devs) as many:
                    assert many.collective == ("host" if no_nccl else "nccl")
                    cnt, stn = many.trace_fluxmap(scenes, altb.source(), n, gm, seed=SEED)
            finally:
                os.environ.pop("ALTB_NO_NCCL", None)
            assert np.array_equal(one, cnt)
            for s in range(2):
                for key in ("n_rays", "n_exited", "n_exit_port", "n_absorbed", "n_suspended", "n_bounces"):
                    assert st1[s][key] == stn[s][key]


def test_map_stage_alone_on_host_records(ctx, oracle, altb):
    """altb_map_records: the map stage on caller-provided records (what a host that already holds
    GetExited()-style results would call), every map mode, ragged batch."""
    sc_g, sc_o = altb.scene(theta_max=166.0), oracle.scene(theta_max=166.0)
    rec, _ = oracle.trace(sc_o, oracle.source(), 25_003, seed=3, prec=oracle.F32)
    ctx.set_batch(9_999)
    try:
        for mode in ("LINE", "TRACEONCE_COMPAT", "DIRECTION"):
            g = ctx.map_records(sc_g, altb.map_spec(mode=getattr(altb, "MAP_" + mode)), rec)
            o = oracle.map_records(sc_o, oracle.map_spec(mode=getattr(oracle, "MAP_" + mode)), rec, prec=oracle.F32)
            assert np.array_equal(g, o), mode
        gm = altb.map_spec(20, 10, 100.0, 40.0, altb.MAP_PER_POSITION, rays_per_position=125)
        om = oracle.map_spec(20, 10, 100.0, 40.0, oracle.MAP_PER_POSITION, rays_per_position=125)
        whole = ctx.map_records(sc_g, gm, rec)
        assert np.array_equal(whole, oracle.map_records(sc_o, om, rec, prec=oracle.F32))
        # a shard that does not start at ray 0 says so (altb_map_records_at): the two halves add up to the whole map
        cut = 12_345
        parts = ctx.map_records(sc_g, gm, rec[:cut]) + ctx.map_records(sc_g, gm, rec[cut:], ray_id0=cut)
        assert np.array_equal(parts, whole) and whole.sum() > 0
    finally:
        ctx.set_batch(0)
    assert ctx.map_records(sc_g, altb.map_spec(mode=altb.MAP_LINE), rec[:0]).sum() == 0


def test_polylines_bit_exact(ctx, oracle, altb):
    """altb_trace_paths: the per-ray hit-point polylines the reference draws with MakePolyLine3D
    (makeIntegratingSphereNRays.C:69-72), small N."""
    for kw, src in ((dict(theta_max=170.0), (-60.0, 0.0, -75.0)),
                    (dict(theta_max=170.0, world_half=200.0, reflectance=1.0, roughness=0.0, max_bounces=10000), (-60.0, 0.0, -80.0))):
        n, mp = 2000, 64
        g_pts, g_np, g_st = ctx.trace_paths(altb.scene(**kw), altb.source(src), n, mp, seed=SEED)
        o_pts, o_np, o_st = oracle.trace_paths(oracle.scene(**kw), oracle.source(src), n, mp, seed=SEED)
        assert np.array_equal(g_np, o_np) and np.array_equal(g_st, o_st)
        assert np.array_equal(g_pts.view(np.uint32), o_pts.view(np.uint32))
        rec, _ = ctx.trace_records(altb.scene(**kw), altb.source(src), n, seed=SEED)
        assert np.array_equal(g_np, 1 + rec["n_hits"] + (rec["status"] == altb.EXITED))      # = ARay::GetNpoints
        assert np.allclose(g_pts[:, 0], np.float32(src)) and np.allclose(g_pts[:, 1], g_pts[0, 1])   # pencil beam: same first hit
        short = g_np <= mp
        last = g_pts[np.flatnonzero(short), g_np[short] - 1]
        assert np.array_equal(last, rec["pos"][short])                                       # = GetLastPoint
        r = np.linalg.norm(g_pts[:, 1].astype(np.float64), axis=1)
        assert np.allclose(r, 100.1, atol=1e-4)                                              # hits lie on the inner sphere


@pytest.mark.parametrize("nt,npb,width", [(1, 1, 40.0), (7, 3, 40.0), (90, 45, 10.0), (181, 91, 40.0), (33, 2, 5.0), (2, 64, 60.0),
                                          (17, 250, 25.0), (250, 9, 80.0), (251, 33, 120.0)])
def test_line_map_odd_grids(ctx, oracle, altb, nt, npb, width):
    """Tile / super-tile culling must stay conservative for any grid shape (partial tiles, single rows, odd sizes); the
    row-stationary rectangle kernel must cope with rectangles of more than 32 columns (column chunks: 250 columns at the
    pole), of more than 32 row pairs (row blocks: an 80 cm detector on 0.36 deg rows) and with a last row pair that has no
    second row (odd n_theta)."""
    n = 12_000
    for mode in ("LINE", "TRACEONCE_COMPAT"):
        g, _ = ctx.trace_fluxmap(altb.scene(theta_max=168.0), altb.source(), n, altb.map_spec(nt, npb, 100.0, width, getattr(altb, "MAP_" + mode)), seed=5)
        o, _ = oracle.fluxmap(oracle.scene(theta_max=168.0), oracle.source(), n, oracle.map_spec(nt, npb, 100.0, width, getattr(oracle, "MAP_" + mode)),
                              seed=5, prec=oracle.F32)
        assert np.array_equal(g[0], o), (mode, int(g.sum()), int(o.sum()))
    d, _ = ctx.trace_fluxmap(altb.scene(theta_max=168.0), altb.source(), n, altb.map_spec(nt, npb, 100.0, width, altb.MAP_DIRECTION), seed=5)
    od, _ = oracle.fluxmap(oracle.scene(theta_max=168.0), oracle.source(), n, oracle.map_spec(nt, npb, 100.0, width, oracle.MAP_DIRECTION), seed=5, prec=oracle.F32)
    assert np.array_equal(d[0], od)


def test_line_map_detector_geometry_variants(ctx, oracle, altb):
    """Other detector radii / widths than the 100 cm / 40 cm of fluxAtObserverFast.C:1276-1277 (e.g. the 10 cm default
    Detector() of fluxAtObserver.C / nonLambertianFlux.C)."""
    n = 15_000
    for radius, width in ((100.0, 10.0), (50.0, 40.0), (250.0, 80.0), (100.0, 199.0)):
        g, _ = ctx.trace_fluxmap(altb.scene(), altb.source(), n, altb.map_spec(60, 30, radius, width, altb.MAP_LINE), seed=8)
        o, _ = oracle.fluxmap(oracle.scene(), oracle.source(), n, oracle.map_spec(60, 30, radius, width, oracle.MAP_LINE), seed=8, prec=oracle.F32)
        assert np.array_equal(g[0], o), (radius, width)
