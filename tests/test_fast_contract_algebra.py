"""CPU check of the three operations the fast contracts leave out of the exact sequence (DESIGN.md section 2,
csrc/altb_math.cuh: ALTB_FAST_FLIP_MIN, ALTB_FAST_SKIP_SETMAG, |x| under the Box-Muller root): each is a no-op up to FP32
rounding.  float32 numpy, same operation order as the kernels; no oracle, no GPU."""
import numpy as np

f32 = np.float32
RNG = np.random.default_rng(20261018)


def unit(n):
    v = RNG.normal(size=(n, 3)).astype(f32)
    return (v / np.sqrt((v * v).sum(axis=1, dtype=f32))[:, None]).astype(f32)


def test_mirror_about_the_true_surface_min_form_equals_the_branch():
    """bounce_step: `if (dn < 0) { d -= 2 dn n; dn = -dn; }` against `d -= 2 min(dn, 0) n; dn = |dn|`."""
    n, d = unit(200_000), unit(200_000)
    dn = (d * n).sum(axis=1, dtype=f32)
    branch = np.where((dn < 0)[:, None], d + (f32(-2.0) * dn)[:, None] * n, d).astype(f32)
    minform = (d + (f32(-2.0) * np.minimum(dn, f32(0.0)))[:, None] * n).astype(f32)
    assert np.array_equal(branch, minform)              # values equal; only a -0 component could come out as +0
    assert np.array_equal(np.where(dn < 0, -dn, dn), np.abs(dn))
    assert (dn < 0).mean() > 0.4                        # the mirrored half is really exercised


def test_reflected_unit_vector_needs_no_setmag():
    """brdf_mix: b = inc - 2 (inc.n) n of two f32 unit vectors; the exact contract applies one Newton step towards |b| = 1
    (reflect.SetMag(1.0), nonLambertianFlux.C:172-176).  |b|^2 - 1 stays within a few ulp, and the sampled direction built
    on b is normalised afterwards: leaving the step out moves the result by less than 4e-7."""
    inc, n = unit(200_000), unit(200_000)
    m = f32(-2.0) * (inc * n).sum(axis=1, dtype=f32)
    b = (inc + m[:, None] * n).astype(f32)
    bb = (b * b).sum(axis=1, dtype=f32)
    assert np.abs(bb.astype(np.float64) - 1.0).max() < 1.5e-6
    sc = (bb * f32(-0.5) + f32(1.5)).astype(f32)
    b_exact = (sc[:, None] * b).astype(f32)
    # direction of a sample c0 o + c1 (b x o) + c2 b does not depend on |b| to first order: compare the unit vectors
    ub = b / np.linalg.norm(b.astype(np.float64), axis=1)[:, None]
    ue = b_exact / np.linalg.norm(b_exact.astype(np.float64), axis=1)[:, None]
    assert np.abs(ub - ue).max() < 4e-7


def test_box_muller_radicand_abs_equals_clamp_except_at_u1_equal_one():
    """box_muller (fast): rad^2 = -2 ln u1 = 27.725887 - 1.3862944 lg2(k), k = 1 .. 2^20.  max(., 0) and |.| differ only where the
    value is negative: nowhere below k = 2^20, and there by at most 2e-6 (radius 1.4e-3 instead of 0, probability 2^-20)."""
    k = np.arange(1, (1 << 20) + 1, dtype=np.float64)
    lg = np.log2(k).astype(f32)                                              # MUFU.LG2: relative error 2^-22, exact at powers of two
    for err in (f32(0.0), f32(2.0 ** -22), f32(-(2.0 ** -22))):              # its error band
        x = (lg * (f32(1.0) + err)).astype(f32) * f32(-1.3862944) + f32(27.725887)
        neg = x < 0
        assert not neg[:-1024].any()
        assert np.abs(x[neg]).max(initial=0.0) < 1.5e-5
    x = lg * f32(-1.3862944) + f32(27.725887)
    assert abs(float(x[-1])) < 2e-6 and (x[:-1] > 0).all()
