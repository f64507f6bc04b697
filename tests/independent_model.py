"""A THIRD implementation of the path, independent of both the CUDA kernels and oracle/ -- test infrastructure only.

Why: the kernels are checked bit for bit against the oracle's F32 mode, which mirrors them operation by operation, and the
oracle's F64 mode shares the state machine (oracle_core.inc).  An error of the ALGORITHM common to both would pass every
such test.  This module restates SURVEY.md appendix A (the reconstruction of ROBAST's AOpticsManager::TraceNonSequential
for the scene of flux_at_observer/fluxAtObserverFast.C:192-230) from the text alone, in vectorised numpy double precision,
with deliberately different means:

* every event is found by a GENERIC nearest-hit search over all solid surfaces (inner sphere, conical port edge, outer
  sphere) -- no closed form for points on the sphere, no port-crossing case analysis, no re-projection (unit vectors are
  re-normalised once per hit);
* rotations are Rodrigues rotations of vectors, not compositions in a local frame; the tangent frame comes from a cross
  product with a coordinate axis, not from the branch-free basis;
* numpy's PCG64 stream and numpy's normal deviates instead of Philox4x32-10 and the Box-Muller pair.

It can therefore only be compared STATISTICALLY (tests/test_independent_model.py): port fraction, status fractions, mean
hit count, port-edge hit rate, exit-direction distribution, line-map profile.
"""
import numpy as np

EXITED, ABSORBED, SUSPENDED = 1, 2, 3


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _rodrigues(v, axis, ang):
    """v rotated about the unit vector `axis` by `ang` (right-handed)."""
    c, s = np.cos(ang)[:, None], np.sin(ang)[:, None]
    return v * c + np.cross(axis, v) * s + axis * (np.sum(axis * v, axis=1, keepdims=True) * (1 - c))


def _perp(n):
    """some unit vector perpendicular to n: cross product with the coordinate axis n is least aligned with."""
    k = np.argmin(np.abs(n), axis=1)
    e = np.zeros_like(n)
    e[np.arange(len(n)), k] = 1.0
    return _unit(np.cross(n, e))


def trace(n_rays, theta_max=170.0, r_inner=100.1, r_outer=101.0, world_half=300.0, reflectance=0.99, roughness=0.01,
          max_bounces=50000, src=(-60.0, 0.0, -75.0), direction=(5.0, 0.0, 0.0), seed=1, tiny=1e-7):
    """Returns dict(status, pos, dir, n_hits, edge_hits): per-ray final status, last point, last direction, surface hits, and
    how many of them were on the conical port edge."""
    rng = np.random.default_rng(seed)
    R1, R2, H = float(r_inner), float(r_outer), float(world_half)
    th = np.deg2rad(theta_max)
    cth, sth = np.cos(th), np.sin(th)
    T2 = (sth / cth) ** 2
    p = np.tile(np.asarray(src, float), (n_rays, 1))
    d = np.tile(_unit(np.asarray(direction, float)), (n_rays, 1))
    status = np.zeros(n_rays, np.int32)
    hits = np.zeros(n_rays, np.int64)
    edge_hits = np.zeros(n_rays, np.int64)
    out_pos = np.zeros((n_rays, 3))
    out_dir = np.zeros((n_rays, 3))
    idx = np.arange(n_rays)

    while idx.size:
        m = idx.size
        best_t = np.full(m, np.inf)
        best_kind = np.zeros(m, np.int8)          # 0 nothing (leaves), 1 inner sphere, 2 cone edge, 3 outer sphere

        def consider(t, ok, kind):
            ok = ok & (t > tiny) & (t < best_t)
            best_t[ok] = t[ok]
            best_kind[ok] = kind

        pd = np.sum(p * d, axis=1)
        pp = np.sum(p * p, axis=1)
        # spheres: |p + t d|^2 = R^2; a hit counts where the shell is solid (polar angle <= theta_max <=> z / R >= cos theta_max)
        for R, kind in ((R1, 1), (R2, 3)):
            disc = pd * pd - (pp - R * R)
            has = disc >= 0
            sq = np.sqrt(np.where(has, disc, 0.0))
            for t in (-pd - sq, -pd + sq):
                z = p[:, 2] + t * d[:, 2]
                consider(t, has & (z >= R * cth), kind)
        # cone x^2 + y^2 = tan^2(theta_max) z^2, lower nappe, between the two radii
        A = d[:, 0] ** 2 + d[:, 1] ** 2 - T2 * d[:, 2] ** 2
        B = p[:, 0] * d[:, 0] + p[:, 1] * d[:, 1] - T2 * p[:, 2] * d[:, 2]
        Cc = p[:, 0] ** 2 + p[:, 1] ** 2 - T2 * p[:, 2] ** 2
        disc = B * B - A * Cc
        has = (disc >= 0) & (np.abs(A) > 1e-300)
        sq = np.sqrt(np.where(has, disc, 0.0))
        Asafe = np.where(has, A, 1.0)
        for t in ((-B - sq) / Asafe, (-B + sq) / Asafe):
            x = p + t[:, None] * d
            r2 = np.sum(x * x, axis=1)
            consider(t, has & (x[:, 2] < 0) & (r2 >= R1 * R1) & (r2 <= R2 * R2), 2)

        # ---- rays that meet nothing leave: world-box point
        gone = best_kind == 0
        if gone.any():
            g = np.flatnonzero(gone)
            with np.errstate(divide="ignore", invalid="ignore"):
                tb = np.where(d[g] > 0, (H - p[g]) / d[g], np.where(d[g] < 0, (-H - p[g]) / d[g], np.inf))
            t = tb.min(axis=1)
            out_pos[idx[g]] = p[g] + t[:, None] * d[g]
            out_dir[idx[g]] = d[g]
            status[idx[g]] = EXITED
        keep = ~gone
        idx, p, d, best_t, best_kind = idx[keep], p[keep], d[keep], best_t[keep], best_kind[keep]
        if not idx.size:
            break
        m = idx.size
        h = p + best_t[:, None] * d
        hits[idx] += 1
        edge_hits[idx] += best_kind == 2
        # facing normal: inner sphere -> toward the centre; outer sphere -> outward; cone -> the polar unit vector at theta_max
        nrm = np.empty_like(h)
        s1 = best_kind == 1
        nrm[s1] = -_unit(h[s1])
        s3 = best_kind == 3
        nrm[s3] = _unit(h[s3])
        ce = best_kind == 2
        if ce.any():
            rho = np.hypot(h[ce, 0], h[ce, 1])
            nrm[ce] = np.stack([cth * h[ce, 0] / rho, cth * h[ce, 1] / rho, np.full(rho.shape, -sth)], axis=1)
        # 1. absorption
        dead = reflectance < rng.random(m)
        # 2. Gaussian roughness: tilt the normal by g = sigma N(0,1) about an axis perpendicular to it at a uniform azimuth
        nt = nrm
        if roughness != 0.0:
            axis = _rodrigues(_perp(nrm), nrm, 2 * np.pi * rng.random(m))
            nt = _rodrigues(nrm, axis, roughness * rng.standard_normal(m))
        # 3. Lambert about the tilted normal: polar angle asin(sqrt(u)) away from it, uniform azimuth about it
        tilt_axis = _perp(nt)
        dn = _rodrigues(nt, tilt_axis, np.arcsin(np.sqrt(rng.random(m))))
        dn = _rodrigues(dn, nt, 2 * np.pi * rng.random(m))
        # a direction that points into the wall is mirrored about the TRUE tangent plane
        into = np.sum(dn * nrm, axis=1)
        flip = into < 0
        dn[flip] -= 2 * into[flip, None] * nrm[flip]
        dn = _unit(dn)      # unit vectors stay unit vectors: without this the rounding error of |d| triples at every bounce
        # bookkeeping: absorbed rays keep their incoming direction, suspended ones the new one
        if dead.any():
            a = np.flatnonzero(dead)
            out_pos[idx[a]] = h[a]; out_dir[idx[a]] = d[a]; status[idx[a]] = ABSORBED
        susp = ~dead & (hits[idx] >= max_bounces)
        if susp.any():
            a = np.flatnonzero(susp)
            out_pos[idx[a]] = h[a]; out_dir[idx[a]] = dn[a]; status[idx[a]] = SUSPENDED
        go = ~dead & ~susp
        idx, p, d = idx[go], h[go], dn[go]
    return dict(status=status, pos=out_pos, dir=out_dir, n_hits=hits, edge_hits=edge_hits)


def line_map(pos, direction, n_theta=180, n_phi=90, det_radius=100.0, det_width=40.0, rows=None):
    """Detector::setPosition + checkIntersection (fluxAtObserverFast.C:61-119) in numpy, for the theta rows `rows`
    (default: all): counts[len(rows), n_phi] of lines (pos, direction) that hit the disk at each position."""
    rows = np.arange(n_theta) if rows is None else np.asarray(rows)
    counts = np.zeros((len(rows), n_phi), np.int64)
    c0 = np.array([0.0, 0.0, -100.0])
    for a, i in enumerate(rows):
        th = np.deg2rad((i + 0.5) * 90.0 / n_theta)
        for j in range(n_phi):
            ph = np.deg2rad((j + 0.5) * 360.0 / n_phi)
            P = c0 + det_radius * np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), -np.cos(th)])
            dd = P - c0
            nrm = np.array([-dd[1], dd[0], dd[2]]) / np.linalg.norm(dd)
            dot = direction @ nrm
            ok = np.abs(dot) >= 1e-10
            t = -((pos - P) @ nrm) / np.where(ok, dot, 1.0)
            I = pos + t[:, None] * direction
            r2 = np.sum(np.cross(np.broadcast_to(nrm, I.shape), I - P) ** 2, axis=1)
            counts[a, j] = np.count_nonzero(ok & (r2 <= (det_width / 2) ** 2))
    return counts
