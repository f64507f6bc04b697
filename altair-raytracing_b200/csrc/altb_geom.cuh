// altb_geom.cuh -- scene constants and the double-precision "slow path" (port crossing, conical
// port edge between r_inner and r_outer, world box), shared by host setup code and the kernels.
// Only IEEE + - * / sqrt, no contraction (-fmad=false / -ffp-contract=off).
// Geometry follows SURVEY.md appendix A.1-A.3 (TGeoSphere(rmin,rmax,0,thetaMax) + TGeoBBox world of
// flux_at_observer/fluxAtObserverFast.C:199-204).
#pragma once
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#include "altb_math.cuh"

#define ALTB_HD __host__ __device__ __forceinline__

namespace altb {

enum { EV_WALL = 1, EV_EDGE = 2, EV_EXIT = 3, EV_OUTER = 4 };   // OUTER: the shell's outer surface hit from outside (brdf_kind 3 only)

struct Geom {
    double R1, R2, R1sq, R2sq, zc, T2, cth, sth, H, exit_z;
    int lambertian, brdf_kind, max_bounces, count_all;
};

struct KConsts {
    float rho, sigma, two_r1, neg_inv_r1, nr_c, zc, p_spec, brdf_s, exit_zf, lobe_ang, inv_r2; int lobe_n;
    int tilt_small, spec_small;   // sigma * max|g| <= 0.9 / brdf_s * max|g| <= 0.9: sin/cos without the quadrant reduction (tilt_small = 2: <= 0.06)
    uint32_t abs_thr, spec_thr;   // integer forms of "rho < u_abs" / "u_sel < p_spec" (altb_math.cuh: HitDraws)
};

// Per-scene constants of a batched launch.  The scenes of ONE launch differ only in theta_max (the port-angle series of
// fluxAtObserverFast.C:1641-1673); everything the hot loop reads from KConsts is common to the launch, the port plane zcf
// rides along in a per-lane register.  The table travels in the kernel's parameter space (no upload, no lifetime): it is
// read with indexed constant loads when a warp claims work and in the slow paths.
struct SceneSlot {
    double zc, T2, cth, sth;           // Geom fields that depend on theta_max
    float zcf;                         // (float) zc
    uint32_t scene;                    // index of the scene in the caller's arrays: counts_base + scene * nb, stats_base + scene * 8
};
static constexpr int MAX_SLOTS = 192;

struct QEntry { float4 a, b; };        // a parked ray: pos.xyz, dir.x | dir.yz, idx, hits

// what the candidate-rectangle computation of the LINE maps needs (altb_kernels.cuh: line_rects)
struct RectParams { int n_theta, n_phi, force_tiles, compat; float det_R, det_Wr; };

struct TraceParams {
    Geom g;               // slot-dependent fields (zc, T2, cth, sth) are those of slot 0; the slow path patches them per ray
    KConsts k;
    int kind0;            // first event of the (identical) source rays
    double x0[3];         // its point
    double d0[3];         // unit source direction
    float x0f[3], d0f[3]; // the same, rounded once (what every fresh ray starts from)
    PhiloxKeys keys;      // round keys of the seed
    uint64_t ray_id0;     // global id of local ray 0 (the launch never straddles a multiple of 2^32: ctr_hi is uniform)
    uint32_t ctr_lo0, ctr_hi;   // the same as two counter words
    uint32_t n;           // rays PER SLOT in this launch
    uint32_t chunk;       // ids a warp claims at a time
    // batched scenes: lane index idx = slot << shift | i, i < n <= 2^shift; work is claimed in chunks q = slot * cps + c
    uint32_t n_slots, shift, imask, cps, n_chunks;
    QEntry* rq;           // resume queues in global memory: [grid * warps][RQCAP]
    unsigned long long* gstat;   // SINK_DIRECTION: per-block statistics [grid][n_slots][STAT_WORDS], zeroed before the launch
    unsigned long long* lane_acc;  // single-slot direction / lines sink: per-THREAD (hits of ended rays, suspended rays) [grid][threads][2], zeroed before the launch
    // SINK_DIRECTION: the maps / statistics the kernel adds to, bins of the direction map
    unsigned long long* counts_base; unsigned long long* stats_base;
    uint32_t nb; int n_theta, n_phi;
    const double* dir_tab;   // bin edges of the direction map (altb_kernels.cuh: direction_bin)
    // SINK_LINES: the escaping rays' test lines go straight to the LINE-map stage -- rectangle list from the front of `lines`,
    // tile list from its back (2 float4 per ray, lines_cap rays in all); n_lines[0] / [1] count them
    RectParams rp; float4* lines; unsigned int* n_lines; uint32_t lines_cap;
    const float2* sincos; // device tables: SC_N sin/cos entries + LG_N log entries (altb_math.cuh: DrawTabs)
    SceneSlot slots[MAX_SLOTS];
};

ALTB_HD void box_exit(const Geom& g, const double* x, const double* d, double* e) {
    double t = INFINITY;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double ti;
        if (d[i] > 0.0) ti = (g.H - x[i]) / d[i];
        else if (d[i] < 0.0) ti = (-g.H - x[i]) / d[i];
        else continue;
        if (ti < t) t = ti;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) e[i] = x[i] + t * d[i];
}

// x0 = inner-sphere crossing inside the opening, heading outward: EDGE(q) or EXIT(e)
ALTB_HD int cap_crossing(const Geom& g, const double* x0, const double* d, double* out) {
    double A = (d[0] * d[0] + d[1] * d[1]) - g.T2 * (d[2] * d[2]);
    double B = (x0[0] * d[0] + x0[1] * d[1]) - g.T2 * (x0[2] * d[2]);
    double C = (x0[0] * x0[0] + x0[1] * x0[1]) - g.T2 * (x0[2] * x0[2]);
    double disc = B * B - A * C;
    double s = 0.0;
    bool have = false;
    if (disc >= 0.0) {
        double sq = sqrt(disc);
        if (B > 0.0) { double den = B + sq; if (den > 0.0) { s = -C / den; have = true; } }
        else if (A > 0.0) { s = (sq - B) / A; have = true; }
    }
    if (have && s > 0.0) {
        double q0 = x0[0] + s * d[0], q1 = x0[1] + s * d[1], q2 = x0[2] + s * d[2];
        if (q2 < 0.0) {
            double r2 = (q0 * q0 + q1 * q1) + q2 * q2;
            if (r2 <= g.R2sq) { out[0] = q0; out[1] = q1; out[2] = q2; return EV_EDGE; }
        }
    }
    box_exit(g, x0, d, out);
    return EV_EXIT;
}

// q on the conical port edge, d heading into the opening: WALL / EDGE / EXIT
ALTB_HD int from_edge(const Geom& g, const double* q, const double* d, double* out) {
    double A = (d[0] * d[0] + d[1] * d[1]) - g.T2 * (d[2] * d[2]);
    double B = (q[0] * d[0] + q[1] * d[1]) - g.T2 * (q[2] * d[2]);
    double s_c = INFINITY, xc0 = 0.0, xc1 = 0.0, xc2 = 0.0;
    if (A > 0.0 && B < 0.0) {
        double s = (-2.0 * B) / A;
        double x0 = q[0] + s * d[0], x1 = q[1] + s * d[1], x2 = q[2] + s * d[2];
        if (x2 < 0.0) {
            double r2 = (x0 * x0 + x1 * x1) + x2 * x2;
            if (r2 >= g.R1sq && r2 <= g.R2sq) { s_c = s; xc0 = x0; xc1 = x1; xc2 = x2; }
        }
    }
    double s_in = INFINITY;
    double b = (q[0] * d[0] + q[1] * d[1]) + q[2] * d[2];
    double c0 = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) - g.R1sq;
    if (b < 0.0) {
        if (c0 > 0.0) { double disc = b * b - c0; if (disc > 0.0) s_in = -b - sqrt(disc); }
        else s_in = 0.0;
    }
    if (s_c < s_in) { out[0] = xc0; out[1] = xc1; out[2] = xc2; return EV_EDGE; }
    if (s_in < INFINITY) {
        double xin[3] = {q[0] + s_in * d[0], q[1] + s_in * d[1], q[2] + s_in * d[2]};
        double bb = (xin[0] * d[0] + xin[1] * d[1]) + xin[2] * d[2];
        double cc = ((xin[0] * xin[0] + xin[1] * xin[1]) + xin[2] * xin[2]) - g.R1sq;
        double disc = bb * bb - cc;
        if (disc < 0.0) disc = 0.0;
        double t = sqrt(disc) - bb;
        double h[3] = {xin[0] + t * d[0], xin[1] + t * d[1], xin[2] + t * d[2]};
        double sc = g.R1 / sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
        h[0] *= sc; h[1] *= sc; h[2] *= sc;
        if (h[2] >= g.zc) { out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; return EV_WALL; }
        return cap_crossing(g, h, d, out);
    }
    box_exit(g, q, d, out);
    return EV_EXIT;
}

// brdf_kind 3 (nonLambertianFlux.C:265-268): the re-scattered ray starts where the primary ray ended -- on the world box,
// OUTSIDE the shell -- with any direction.  First event: the solid outer surface S2 (OUTER), or, through the opening of S2,
// the conical port edge (EDGE) or -- across the cavity -- the inner wall (WALL); otherwise it leaves (EXIT, out = world-box
// point; a ray that starts on the box heading outward ends where it starts).
ALTB_HD int from_outside(const Geom& g, const double* p, const double* d, double* out) {
    const double b = (p[0] * d[0] + p[1] * d[1]) + p[2] * d[2];
    const double c2 = ((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]) - g.R2sq;
    if (b < 0.0 && c2 > 0.0) {
        const double disc = b * b - c2;
        if (disc > 0.0) {
            const double s2 = -b - sqrt(disc);
            const double x[3] = {p[0] + s2 * d[0], p[1] + s2 * d[1], p[2] + s2 * d[2]};
            if (x[2] >= g.R2 * g.cth) { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; return EV_OUTER; }
            // x is inside the opening cone: first crossing of the cone (cf. cap_crossing) ...
            const double A = (d[0] * d[0] + d[1] * d[1]) - g.T2 * (d[2] * d[2]);
            const double B = (x[0] * d[0] + x[1] * d[1]) - g.T2 * (x[2] * d[2]);
            const double C = (x[0] * x[0] + x[1] * x[1]) - g.T2 * (x[2] * x[2]);
            const double dc = B * B - A * C;
            double s_c = INFINITY, q0 = 0.0, q1 = 0.0, q2 = 0.0;
            if (dc >= 0.0) {
                const double sq = sqrt(dc);
                double s = 0.0;
                bool have = false;
                if (B > 0.0) { const double den = B + sq; if (den > 0.0) { s = -C / den; have = true; } }
                else if (A > 0.0) { s = (sq - B) / A; have = true; }
                if (have && s > 0.0) {
                    q0 = x[0] + s * d[0]; q1 = x[1] + s * d[1]; q2 = x[2] + s * d[2];
                    if (q2 < 0.0) {
                        const double r2 = (q0 * q0 + q1 * q1) + q2 * q2;
                        if (r2 >= g.R1sq && r2 <= g.R2sq) s_c = s;
                    }
                }
            }
            // ... against the entry into the cavity through the cap of S1
            double s_in = INFINITY;
            const double b1 = (x[0] * d[0] + x[1] * d[1]) + x[2] * d[2];
            const double c1 = ((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]) - g.R1sq;
            if (b1 < 0.0) { const double d1 = b1 * b1 - c1; if (d1 > 0.0) s_in = -b1 - sqrt(d1); }
            if (s_c < s_in) { out[0] = q0; out[1] = q1; out[2] = q2; return EV_EDGE; }
            if (s_in < INFINITY) {
                const double xin[3] = {x[0] + s_in * d[0], x[1] + s_in * d[1], x[2] + s_in * d[2]};
                const double bb = (xin[0] * d[0] + xin[1] * d[1]) + xin[2] * d[2];
                const double cc = ((xin[0] * xin[0] + xin[1] * xin[1]) + xin[2] * xin[2]) - g.R1sq;
                double dd = bb * bb - cc;
                if (dd < 0.0) dd = 0.0;
                const double t = sqrt(dd) - bb;
                double h[3] = {xin[0] + t * d[0], xin[1] + t * d[1], xin[2] + t * d[2]};
                const double sc = g.R1 / sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
                h[0] *= sc; h[1] *= sc; h[2] *= sc;
                if (h[2] >= g.zc) { out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; return EV_WALL; }
                return cap_crossing(g, h, d, out);
            }
            box_exit(g, x, d, out);
            return EV_EXIT;
        }
    }
    box_exit(g, p, d, out);
    return EV_EXIT;
}

// normal of the conical edge at q, pointing into the opening (theta-hat at theta_max)
ALTB_HD void edge_normal(const Geom& g, const double* q, double* n) {
    double rho = sqrt(q[0] * q[0] + q[1] * q[1]);
    if (rho > 0.0) { n[0] = g.cth * (q[0] / rho); n[1] = g.cth * (q[1] / rho); }
    else { n[0] = 0.0; n[1] = 0.0; }
    n[2] = -g.sth;
}

// first event of a ray launched at p0 (inside the cavity) along dir (appendix A.2); <0 on error
ALTB_HD int launch_ray(const Geom& g, const double* p0, const double* dir, double* d0, double* out) {
    double m = sqrt((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]);
    if (!(m > 0.0)) return -1;
#pragma unroll
    for (int i = 0; i < 3; i++) d0[i] = dir[i] / m;
    double b = (p0[0] * d0[0] + p0[1] * d0[1]) + p0[2] * d0[2];
    double c0 = ((p0[0] * p0[0] + p0[1] * p0[1]) + p0[2] * p0[2]) - g.R1sq;
    if (!(c0 < 0.0)) return -1;
    double t = sqrt(b * b - c0) - b;
    double h[3] = {p0[0] + t * d0[0], p0[1] + t * d0[1], p0[2] + t * d0[2]};
    double sc = g.R1 / sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
    h[0] *= sc; h[1] *= sc; h[2] *= sc;
    if (h[2] >= g.zc) { out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; return EV_WALL; }
    return cap_crossing(g, h, d0, out);
}

}  // namespace altb
