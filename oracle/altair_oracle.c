/*
 * altair_oracle.c -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 * See altair_oracle.h for scope, reference citations and the "parity unpinned" statement.
 * Build: see oracle/Makefile (-ffp-contract=off is REQUIRED: every fused multiply-add in
 * the arithmetic contract is written explicitly).
 */
#define _GNU_SOURCE
#include "altair_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EV_WALL 1
#define EV_EDGE 2
#define EV_EXIT 3
#define EV_OUTER 4      /* brdf_kind 3 only: the solid OUTER surface of the shell, hit from outside */
#define PI_D 3.14159265358979323846

/* ------------------------------------------------------------------ scene constants */
typedef struct {
    double R1, R2, R1sq, R2sq, zc, T2, cth, sth, H, exit_z;
    int lambertian, brdf_kind, max_bounces, count_all;
    int lobe_n;           /* brdf_kind 2: integer exponent of the cos^n lobe */
} geom;

typedef struct { float rho, sigma, two_r1, neg_inv_r1, nr_c, zc, p_spec, brdf_s, lobe_ang, inv_r2; } consts_f;
typedef struct { double rho, sigma, two_r1, neg_inv_r1, nr_c, zc, p_spec, brdf_s, lobe_ang, inv_r2; } consts_d;

static int make_geom(const orc_scene* sc, geom* g, consts_f* kf, consts_d* kd) {
    if (!(sc->r_inner > 0) || !(sc->r_outer >= sc->r_inner) || !(sc->world_half > sc->r_outer)) return -1;
    if (!(sc->theta_max_deg > 90.0) || !(sc->theta_max_deg < 180.0)) return -1;
    if (sc->max_bounces < 1) return -1;
    double th = sc->theta_max_deg * PI_D / 180.0;
    g->R1 = sc->r_inner; g->R2 = sc->r_outer;
    g->R1sq = g->R1 * g->R1; g->R2sq = g->R2 * g->R2;
    g->cth = cos(th); g->sth = sin(th);
    g->zc = g->R1 * g->cth;
    double t = g->sth / g->cth;
    g->T2 = t * t;
    g->H = sc->world_half; g->exit_z = sc->exit_z;
    g->lambertian = sc->lambertian; g->brdf_kind = sc->brdf_kind;
    g->max_bounces = sc->max_bounces; g->count_all = sc->count_all_status;
    double ps = 0.0, bs = 0.0;
    if (sc->brdf_kind == 1 || sc->brdf_kind == 3) {   /* nonLambertianFlux.C:156-159: normalise (spec,diff) */
        double sum = sc->brdf_param[1] + sc->brdf_param[2];
        if (!(sum > 0)) return -1;
        ps = sc->brdf_param[1] / sum;
        bs = sc->brdf_param[0] * PI_D / 6.0;
    } else if (sc->brdf_kind == 2) {   /* cos^n lobe of 'nonLambertianFlux copy.C':31-70: exponent, max angle [deg] */
        double ne = sc->brdf_param[0];
        if (!(ne >= 1.0 && ne <= 8.0) || ne != floor(ne)) return -1;
        if (!(sc->brdf_param[1] > 0.0 && sc->brdf_param[1] <= 90.0)) return -1;
        g->lobe_n = (int)ne;
    } else if (sc->brdf_kind != 0) return -1;
    kd->lobe_ang = sc->brdf_kind == 2 ? sc->brdf_param[1] * PI_D / 180.0 : 0.0;
    kf->lobe_ang = (float)kd->lobe_ang;
    kd->rho = sc->reflectance; kd->sigma = sc->roughness_rad;
    kd->two_r1 = 2.0 * g->R1; kd->neg_inv_r1 = -1.0 / g->R1; kd->nr_c = -0.5 / g->R1sq;
    kd->zc = g->zc; kd->p_spec = ps; kd->brdf_s = bs;
    kf->rho = (float)kd->rho; kf->sigma = (float)kd->sigma; kf->two_r1 = (float)kd->two_r1;
    kf->neg_inv_r1 = (float)kd->neg_inv_r1; kf->nr_c = (float)kd->nr_c; kf->zc = (float)kd->zc;
    kf->p_spec = (float)kd->p_spec; kf->brdf_s = (float)kd->brdf_s;
    kd->inv_r2 = 1.0 / g->R2; kf->inv_r2 = (float)kd->inv_r2;
    return 0;
}

/* ------------------------------------------------------------------ double slow path
 * (port crossing, conical port edge, world box) -- IEEE + - * / sqrt only, no contraction,
 * so the kernels' double slow path reproduces it bit for bit.  SURVEY.md appendix A.1/A.3. */
static void box_exit(const geom* g, const double* x, const double* d, double* e) {
    double t = INFINITY;
    for (int i = 0; i < 3; i++) {
        double ti;
        if (d[i] > 0.0) ti = (g->H - x[i]) / d[i];
        else if (d[i] < 0.0) ti = (-g->H - x[i]) / d[i];
        else continue;
        if (ti < t) t = ti;
    }
    for (int i = 0; i < 3; i++) e[i] = x[i] + t * d[i];
}

/* x0 = S1 crossing inside the opening, heading outward.  EDGE(q) or EXIT(e). */
static int cap_crossing(const geom* g, const double* x0, const double* d, double* out) {
    double A = (d[0] * d[0] + d[1] * d[1]) - g->T2 * (d[2] * d[2]);
    double B = (x0[0] * d[0] + x0[1] * d[1]) - g->T2 * (x0[2] * d[2]);
    double C = (x0[0] * x0[0] + x0[1] * x0[1]) - g->T2 * (x0[2] * x0[2]);
    double disc = B * B - A * C;
    double s = 0.0;
    int have = 0;
    if (disc >= 0.0) {
        double sq = sqrt(disc);
        if (B > 0.0) { double den = B + sq; if (den > 0.0) { s = -C / den; have = 1; } }
        else if (A > 0.0) { s = (sq - B) / A; have = 1; }
    }
    if (have && s > 0.0) {
        double q[3] = {x0[0] + s * d[0], x0[1] + s * d[1], x0[2] + s * d[2]};
        if (q[2] < 0.0) {
            double r2 = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
            if (r2 <= g->R2sq) { out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; return EV_EDGE; }
        }
    }
    box_exit(g, x0, d, out);
    return EV_EXIT;
}

/* q on the conical port edge, d heading into the opening.  WALL / EDGE / EXIT. */
static int from_edge(const geom* g, const double* q, const double* d, double* out) {
    double A = (d[0] * d[0] + d[1] * d[1]) - g->T2 * (d[2] * d[2]);
    double B = (q[0] * d[0] + q[1] * d[1]) - g->T2 * (q[2] * d[2]);
    double s_c = INFINITY, xc[3] = {0, 0, 0};
    if (A > 0.0 && B < 0.0) {
        double s = (-2.0 * B) / A;
        double x[3] = {q[0] + s * d[0], q[1] + s * d[1], q[2] + s * d[2]};
        if (x[2] < 0.0) {
            double r2 = (x[0] * x[0] + x[1] * x[1]) + x[2] * x[2];
            if (r2 >= g->R1sq && r2 <= g->R2sq) { s_c = s; xc[0] = x[0]; xc[1] = x[1]; xc[2] = x[2]; }
        }
    }
    double s_in = INFINITY;
    double b = (q[0] * d[0] + q[1] * d[1]) + q[2] * d[2];
    double c0 = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) - g->R1sq;
    if (b < 0.0) {
        if (c0 > 0.0) { double disc = b * b - c0; if (disc > 0.0) s_in = -b - sqrt(disc); }
        else s_in = 0.0;
    }
    if (s_c < s_in) { out[0] = xc[0]; out[1] = xc[1]; out[2] = xc[2]; return EV_EDGE; }
    if (s_in < INFINITY) {
        double xin[3] = {q[0] + s_in * d[0], q[1] + s_in * d[1], q[2] + s_in * d[2]};
        double bb = (xin[0] * d[0] + xin[1] * d[1]) + xin[2] * d[2];
        double cc = ((xin[0] * xin[0] + xin[1] * xin[1]) + xin[2] * xin[2]) - g->R1sq;
        double disc = bb * bb - cc;
        if (disc < 0.0) disc = 0.0;
        double t = sqrt(disc) - bb;
        double h[3] = {xin[0] + t * d[0], xin[1] + t * d[1], xin[2] + t * d[2]};
        double sc = g->R1 / sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
        h[0] *= sc; h[1] *= sc; h[2] *= sc;
        if (h[2] >= g->zc) { out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; return EV_WALL; }
        return cap_crossing(g, h, d, out);
    }
    box_exit(g, q, d, out);
    return EV_EXIT;
}

/* Normal of the conical edge at q, pointing into the opening (theta-hat at theta_max). */
static void edge_normal(const geom* g, const double* q, double* n) {
    double rho = sqrt(q[0] * q[0] + q[1] * q[1]);
    if (rho > 0.0) { n[0] = g->cth * (q[0] / rho); n[1] = g->cth * (q[1] / rho); }
    else { n[0] = 0.0; n[1] = 0.0; }
    n[2] = -g->sth;
}

/* First event of a ray launched at p0 (inside the cavity) along dir (appendix A.2). */
static int launch(const geom* g, const double* p0, const double* dir, double* d0, double* out) {
    double m = sqrt((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]);
    if (!(m > 0.0)) return -1;
    for (int i = 0; i < 3; i++) d0[i] = dir[i] / m;
    double b = (p0[0] * d0[0] + p0[1] * d0[1]) + p0[2] * d0[2];
    double c0 = ((p0[0] * p0[0] + p0[1] * p0[1]) + p0[2] * p0[2]) - g->R1sq;
    if (!(c0 < 0.0)) return -1;
    double t = sqrt(b * b - c0) - b;
    double h[3] = {p0[0] + t * d0[0], p0[1] + t * d0[1], p0[2] + t * d0[2]};
    double sc = g->R1 / sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
    h[0] *= sc; h[1] *= sc; h[2] *= sc;
    if (h[2] >= g->zc) { out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; return EV_WALL; }
    return cap_crossing(g, h, d0, out);
}

/* brdf_kind 3 (nonLambertianFlux.C:265-268): the re-scattered ray starts where the primary ray ended -- on the world box,
 * OUTSIDE the shell -- with any direction.  Its first event: the solid outer surface S2 (EV_OUTER), or, through the opening
 * of S2, the conical port edge (EV_EDGE) or -- across the cavity -- the inner wall (EV_WALL); otherwise it leaves (EV_EXIT,
 * out = world-box point; a ray that starts on the box heading outward ends where it starts). */
static int from_outside(const geom* g, const double* p, const double* d, double* out) {
    double b = (p[0] * d[0] + p[1] * d[1]) + p[2] * d[2];
    double c2 = ((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]) - g->R2sq;
    if (b < 0.0 && c2 > 0.0) {
        double disc = b * b - c2;
        if (disc > 0.0) {
            double s2 = -b - sqrt(disc);
            double x[3] = {p[0] + s2 * d[0], p[1] + s2 * d[1], p[2] + s2 * d[2]};
            if (x[2] >= g->R2 * g->cth) { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; return EV_OUTER; }
            /* x is inside the opening cone: first crossing of the cone (cf. cap_crossing) ... */
            double A = (d[0] * d[0] + d[1] * d[1]) - g->T2 * (d[2] * d[2]);
            double B = (x[0] * d[0] + x[1] * d[1]) - g->T2 * (x[2] * d[2]);
            double C = (x[0] * x[0] + x[1] * x[1]) - g->T2 * (x[2] * x[2]);
            double dc = B * B - A * C;
            double s_c = INFINITY, q[3] = {0, 0, 0};
            if (dc >= 0.0) {
                double sq = sqrt(dc), s = 0.0;
                int have = 0;
                if (B > 0.0) { double den = B + sq; if (den > 0.0) { s = -C / den; have = 1; } }
                else if (A > 0.0) { s = (sq - B) / A; have = 1; }
                if (have && s > 0.0) {
                    q[0] = x[0] + s * d[0]; q[1] = x[1] + s * d[1]; q[2] = x[2] + s * d[2];
                    if (q[2] < 0.0) {
                        double r2 = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
                        if (r2 >= g->R1sq && r2 <= g->R2sq) s_c = s;
                    }
                }
            }
            /* ... against the entry into the cavity through the cap of S1 */
            double s_in = INFINITY;
            double b1 = (x[0] * d[0] + x[1] * d[1]) + x[2] * d[2];
            double c1 = ((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]) - g->R1sq;
            if (b1 < 0.0) { double d1 = b1 * b1 - c1; if (d1 > 0.0) s_in = -b1 - sqrt(d1); }
            if (s_c < s_in) { out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; return EV_EDGE; }
            if (s_in < INFINITY) {
                double xin[3] = {x[0] + s_in * d[0], x[1] + s_in * d[1], x[2] + s_in * d[2]};
                double bb = (xin[0] * d[0] + xin[1] * d[1]) + xin[2] * d[2];
                double cc = ((xin[0] * xin[0] + xin[1] * xin[1]) + xin[2] * xin[2]) - g->R1sq;
                double dd = bb * bb - cc;
                if (dd < 0.0) dd = 0.0;
                double t = sqrt(dd) - bb;
                double h[3] = {xin[0] + t * d[0], xin[1] + t * d[1], xin[2] + t * d[2]};
                double sc = g->R1 / sqrt((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]);
                h[0] *= sc; h[1] *= sc; h[2] *= sc;
                if (h[2] >= g->zc) { out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; return EV_WALL; }
                return cap_crossing(g, h, d, out);
            }
            box_exit(g, x, d, out);
            return EV_EXIT;
        }
    }
    box_exit(g, p, d, out);
    return EV_EXIT;
}

/* ------------------------------------------------------------------ f32 primitives */
static inline float as_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t as_u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline void sincos_poly_f32(float x, int q, float* s, float* c) {
    float x2 = x * x;
    float ps = fmaf(x2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, x2, -1.6666654611e-1f);
    float sn = fmaf(x * x2, ps, x);
    float pc = fmaf(x2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, x2, 4.166664568298827e-2f);
    float cs = fmaf(x2 * x2, pc, fmaf(x2, -0.5f, 1.0f));
    switch (q & 3) {
        case 0: *s = sn;  *c = cs;  break;
        case 1: *s = cs;  *c = -sn; break;
        case 2: *s = -sn; *c = -cs; break;
        default: *s = -cs; *c = sn; break;
    }
}

void orc_sincos2pi_f32(float u, float* s, float* c) {
    float q = rintf(u * 4.0f);
    float r = fmaf(q, -0.25f, u);
    sincos_poly_f32(r * 6.2831855f, (int)q, s, c);
}

/* Azimuths are fixed-point turn fractions cut from the Philox block, so the kernels look their sin/cos up:
 * tab[i] = orc_sincos2pi_f32(i / 8192) (8192 entries), and a 20-bit fraction q = hi:13 | lo:7 adds the second-order
 * rotation by B = 2 pi lo / 2^20 (csrc/altb_math.cuh: SinCosTab).  The oracle restates exactly that. */
static float g_sc_tab[8192][2];
__attribute__((constructor)) static void sc_tab_init(void) {
    for (int i = 0; i < 8192; i++) orc_sincos2pi_f32((float)i * 0x1p-13f, &g_sc_tab[i][0], &g_sc_tab[i][1]);
}
void orc_sincos2pi_q13(uint32_t q, float* s, float* c) { *s = g_sc_tab[q & 8191u][0]; *c = g_sc_tab[q & 8191u][1]; }
void orc_sincos2pi_q20(uint32_t q, float* s, float* c) {
    const float* a = g_sc_tab[(q >> 7) & 8191u];
    float B = (float)(q & 127u) * (6.2831855f * 0x1p-20f);
    float h = -0.5f * B;
    *s = fmaf(fmaf(h, a[0], a[1]), B, a[0]);
    *c = fmaf(fmaf(h, a[1], -a[0]), B, a[1]);
}
/* tape records carry the fractions as floats.  g_full_az (orc_replay_ex, ORC_REPLAY_FULL_AZIMUTH): a tape recorded elsewhere
 * holds arbitrary uniforms -- sin / cos of 2 pi u at full float precision (the polynomial), as altb_replay_ex does; set
 * before the parallel region, read-only inside. */
static int g_full_az = 0;
static inline void sincos2pi_u13(float u, float* s, float* c) {
    if (g_full_az) orc_sincos2pi_f32(u, s, c); else orc_sincos2pi_q13((uint32_t)(u * 8192.0f), s, c);
}
static inline void sincos2pi_u20(float u, float* s, float* c) {
    if (g_full_az) orc_sincos2pi_f32(u, s, c); else orc_sincos2pi_q20((uint32_t)(u * 1048576.0f) & 0xfffffu, s, c);
}

/* Contract: |x| <= 0.9 is evaluated directly (quadrant 0 polynomials), anything larger is reduced first. */
void orc_sincos_f32(float x, float* s, float* c) {
    if (fabsf(x) <= 0.9f) { sincos_poly_f32(x, 0, s, c); return; }
    float q = rintf(x * 0.63661975f);
    float r = fmaf(q, -1.5707964f, x);
    r = fmaf(q, 4.3711388e-8f, r);
    sincos_poly_f32(r, (int)q, s, c);
}

/* ln(k 2^-20) for the 20-bit Box-Muller integer k = 1 .. 2^20 (csrc/altb_math.cuh: DrawTabs::log_u20): table over the
 * top 7 mantissa bits of (float)k -- lg[hi] = (ln(m_hi'), 1/m_hi, e_adj), m_hi = 1 + hi/128, m_hi' = m_hi or m_hi/2
 * (hi >= 53) -- plus a degree-4 log1p of the remainder. */
static float g_lg_tab[128][3];
__attribute__((constructor)) static void lg_tab_init(void) {
    for (int hi = 0; hi < 128; hi++) {
        double mh = 1.0 + hi / 128.0;
        int half = hi >= 53;
        g_lg_tab[hi][0] = (float)log(half ? mh * 0.5 : mh);
        g_lg_tab[hi][1] = (float)(1.0 / mh);
        g_lg_tab[hi][2] = half ? -146.0f : -147.0f;
    }
}
float orc_log_u20(uint32_t k) {
    uint32_t b = as_u((float)k);
    const float* t = g_lg_tab[(b >> 16) & 0x7fu];
    float base = fmaf((float)(b >> 23) + t[2], 0.69314718f, t[0]);
    float mf = as_f((b & 0x007fffffu) | 0x3f800000u);
    float mh = as_f((b & 0x007f0000u) | 0x3f800000u);
    float r = (mf - mh) * t[1];
    float q = fmaf(r, -0.25f, 0.33333334f);
    q = fmaf(q, r, -0.5f);
    q = fmaf(q, r, 1.0f);
    return fmaf(q, r, base);
}

static inline void sincos2pi_d(double u, double* s, double* c) { double x = 2.0 * PI_D * u; *s = sin(x); *c = cos(x); }
static inline void sincos_d(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }

/* ------------------------------------------------------------------ Philox + draws */
/* rounds of the generator behind orc_draws / orc_trace: 10 (default; contracts EXACT and FAST of the library) or 7
 * (ALTB_CONTRACT_FAST7: the fewest rounds that pass BigCrush, Salmon et al. 2011 table 2).  Process-wide, test infrastructure. */
static int g_philox_rounds = 10;
int orc_set_philox_rounds(int rounds) {
    if (rounds != 7 && rounds != 10) return -1;
    g_philox_rounds = rounds;
    return 0;
}
static void philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], int rounds);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_r(ctr, key, out, 10); }
void orc_philox4x32_7(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_r(ctr, key, out, 7); }
static void philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], int rounds) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ONE Philox4x32-10 block (128 bits) per surface hit, counter = (ray_id lo, ray_id hi, k, 0), key = seed:
 *   w0: u_abs 24 b | 8 b -> u_sel (low byte)      w1: u_r 24 b | 8 b -> bm_u1 (low byte)
 *   w2: u_phi 20 b | 12 b -> bm_u1 (high bits)    w3: u_psi 13 b | bm_u2 13 b | 6 b -> u_sel (high bits)
 * (g0, g1) = Box-Muller of (bm_u1 in (0,1] with 20 bits, bm_u2 with 13 bits). */
static void draws_blk(uint64_t seed, uint64_t ray_id, uint32_t k, uint32_t blk, float out[ORC_DRAWS_PER_HIT]);
void orc_draws(uint64_t seed, uint64_t ray_id, uint32_t k, float out[ORC_DRAWS_PER_HIT]) { draws_blk(seed, ray_id, k, 0u, out); }
/* counter word 3: 0 = the surface hits of the (primary) trace; 1, 2 = the cos^n rejection loop (orc_draws_lobe);
 * 4 = the post-hoc re-scatter of brdf_kind 3 (k = 0), 5 = the surface hits of its second trace */
static void draws_blk(uint64_t seed, uint64_t ray_id, uint32_t k, uint32_t blk, float out[ORC_DRAWS_PER_HIT]) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)ray_id, (uint32_t)(ray_id >> 32), k, blk};
    uint32_t w[4];
    philox4x32_r(ctr, key, w, g_philox_rounds);
    out[0] = (float)(w[0] >> 8) * 0x1p-24f;
    out[1] = (float)(w[1] >> 8) * 0x1p-24f;
    out[2] = (float)(w[2] >> 12) * 0x1p-20f;
    out[3] = (float)(((w[3] & 0x3fu) << 8) | (w[0] & 0xffu)) * 0x1p-14f;
    out[4] = (float)(w[3] >> 19) * 0x1p-13f;
    uint32_t t = ((w[2] & 0xfffu) << 8) | (w[1] & 0xffu);
    float rad = sqrtf(2.0f * fabsf(orc_log_u20(t + 1u)));   /* u1 = (t+1) 2^-20 in (0,1]; log <= 0, |.| keeps u1 = 1 at +0 */
    float s, c;
    orc_sincos2pi_q13((w[3] >> 6) & 0x1fffu, &s, &c);
    out[5] = rad * c; out[6] = rad * s;
    out[7] = 0.0f;
}

/* brdf_kind 2: the rejection loop of generateScatteredDirection ('nonLambertianFlux copy.C':47-69) only decides
 * the polar angle (acceptance cos^n(theta) does not depend on phi), so it lives on the RNG side: slot [1] of the draw
 * record becomes the ACCEPTED r1 (theta = max_angle * r1).  Attempts come from extra Philox blocks
 * (counter word3 = 1, 2), four (r1, r3) pairs of 16 + 16 bits per block; after 8 rejections (p < 1e-4) the last r1 stands. */
void orc_draws_lobe(uint64_t seed, uint64_t ray_id, uint32_t k, int lobe_n, float lobe_ang, float out[ORC_DRAWS_PER_HIT]) {
    orc_draws(seed, ray_id, k, out);
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    float r1 = 0.0f;
    for (uint32_t blk = 1; blk <= 2; blk++) {
        uint32_t ctr[4] = {(uint32_t)ray_id, (uint32_t)(ray_id >> 32), k, blk}, w[4];
        orc_philox4x32_10(ctr, key, w);
        for (int a = 0; a < 4; a++) {
            r1 = (float)(w[a] >> 16) * 0x1p-16f;
            float r3 = (float)(w[a] & 0xffffu) * 0x1p-16f;
            float s, c, p = 1.0f;
            orc_sincos_f32(lobe_ang * r1, &s, &c);
            for (int e = 0; e < lobe_n; e++) p = p * c;
            if (r3 <= p) { out[1] = r1; return; }
        }
    }
    out[1] = r1;
}

/* ------------------------------------------------------------------ the two instantiations */
#define REAL float
#define SUF(n) n##_f
#define RC(x) x##f
#define FMA(a, b, c) fmaf(a, b, c)
#define SQRT(x) sqrtf(x)
#define FABS(x) fabsf(x)
#define COPYSIGN(a, b) copysignf(a, b)
#define SINCOS2PI_13(u, s, c) sincos2pi_u13(u, s, c)
#define SINCOS2PI_20(u, s, c) sincos2pi_u20(u, s, c)
#define SINCOS(x, s, c) orc_sincos_f32(x, s, c)
#include "oracle_core.inc"
#undef REAL
#undef SUF
#undef RC
#undef FMA
#undef SQRT
#undef FABS
#undef COPYSIGN
#undef SINCOS2PI_13
#undef SINCOS2PI_20
#undef SINCOS

#define REAL double
#define SUF(n) n##_d
#define RC(x) x
#define FMA(a, b, c) ((a) * (b) + (c))
#define SQRT(x) sqrt(x)
#define FABS(x) fabs(x)
#define COPYSIGN(a, b) copysign(a, b)
#define SINCOS2PI_13(u, s, c) sincos2pi_d(u, s, c)
#define SINCOS2PI_20(u, s, c) sincos2pi_d(u, s, c)
#define SINCOS(x, s, c) sincos_d(x, s, c)
#include "oracle_core.inc"
#undef REAL
#undef SUF
#undef RC
#undef FMA
#undef SQRT
#undef FABS
#undef COPYSIGN
#undef SINCOS2PI_13
#undef SINCOS2PI_20
#undef SINCOS

/* ------------------------------------------------------------------ per-ray drivers */
typedef struct { const float* tape; uint64_t n_rec; } tape_src;   /* tape == NULL -> Philox */

typedef struct { double pos[3], dir[3]; uint32_t n_hits; uint32_t status; } result_d;

__attribute__((target_clones("arch=haswell","default")))
static void run_ray(const geom* g, const consts_f* kf, const consts_d* kd, int prec,
                    int kind0, const double* x0, const double* d0,
                    uint64_t seed, uint64_t ray_id, const tape_src* ts,
                    orc_record* rec, result_d* rd, float* tape_out, uint64_t* tape_n) {
    float dr[ORC_DRAWS_PER_HIT];
    int st;
    uint32_t k = 0;
    if (prec == ORC_F32) {
        state_f s;
        st = start_f(g, kf, &s, kind0, x0, d0);
        while (!st) {
            if (ts && ts->tape) {
                if (k >= ts->n_rec) { st = ORC_TAPE_END; break; }
                memcpy(dr, ts->tape + 8 * (uint64_t)k, sizeof dr);
            } else if (g->brdf_kind == 2) orc_draws_lobe(seed, ray_id, k, g->lobe_n, kf->lobe_ang, dr);
            else orc_draws(seed, ray_id, k, dr);
            if (tape_out) memcpy(tape_out + 8 * (uint64_t)k, dr, sizeof dr);
            k++;
            st = bounce_f(g, kf, &s, dr);
        }
        if (g->brdf_kind == 3 && st == ORC_EXITED && !(ts && ts->tape)) {     /* post-hoc re-scatter + second trace */
            const uint32_t h1 = s.n_hits;
            draws_blk(seed, ray_id, 0u, 4u, dr);
            st = rescatter_f(g, kf, &s, d0, dr);
            while (!st) { draws_blk(seed, ray_id, s.n_hits, 5u, dr); st = bounce_f(g, kf, &s, dr); }
            s.n_hits += h1;
        }
        if (tape_n) *tape_n = k;
        if (rec) {
            for (int i = 0; i < 3; i++) { rec->pos[i] = s.pos[i]; rec->dir[i] = s.dir[i]; }
            rec->n_hits = s.n_hits; rec->status = (uint32_t)st;
        }
        if (rd) {
            for (int i = 0; i < 3; i++) { rd->pos[i] = s.pos[i]; rd->dir[i] = s.dir[i]; }
            rd->n_hits = s.n_hits; rd->status = (uint32_t)st;
        }
    } else {
        state_d s;
        st = start_d(g, kd, &s, kind0, x0, d0);
        while (!st) {
            if (ts && ts->tape) {
                if (k >= ts->n_rec) { st = ORC_TAPE_END; break; }
                memcpy(dr, ts->tape + 8 * (uint64_t)k, sizeof dr);
            } else if (g->brdf_kind == 2) orc_draws_lobe(seed, ray_id, k, g->lobe_n, kf->lobe_ang, dr);
            else orc_draws(seed, ray_id, k, dr);
            k++;
            st = bounce_d(g, kd, &s, dr);
        }
        if (g->brdf_kind == 3 && st == ORC_EXITED && !(ts && ts->tape)) {
            const uint32_t h1 = s.n_hits;
            draws_blk(seed, ray_id, 0u, 4u, dr);
            st = rescatter_d(g, kd, &s, d0, dr);
            while (!st) { draws_blk(seed, ray_id, s.n_hits, 5u, dr); st = bounce_d(g, kd, &s, dr); }
            s.n_hits += h1;
        }
        if (tape_n) *tape_n = k;
        if (rec) {
            for (int i = 0; i < 3; i++) { rec->pos[i] = (float)s.pos[i]; rec->dir[i] = (float)s.dir[i]; }
            rec->n_hits = s.n_hits; rec->status = (uint32_t)st;
        }
        if (rd) {
            for (int i = 0; i < 3; i++) { rd->pos[i] = s.pos[i]; rd->dir[i] = s.dir[i]; }
            rd->n_hits = s.n_hits; rd->status = (uint32_t)st;
        }
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int pick_threads(int n_threads) {
#ifdef _OPENMP
    return n_threads <= 0 ? omp_get_max_threads() : n_threads;
#else
    (void)n_threads; return 1;
#endif
}

static int port_flag_g(const geom* g, const float* pos, uint32_t status) {
    if (!(g->count_all || status == ORC_EXITED)) return 0;
    return pos[2] < (float)g->exit_z;
}

int orc_port_flag(const orc_scene* sc, const orc_record* r) {
    geom g; consts_f kf; consts_d kd;
    if (make_geom(sc, &g, &kf, &kd)) return -1;
    return port_flag_g(&g, r->pos, r->status);
}

static void add_stats(const geom* g, const orc_record* r, orc_stats* s) {
    s->n_rays += 1; s->n_bounces += r->n_hits;
    if (r->status == ORC_EXITED) s->n_exited += 1;
    else if (r->status == ORC_ABSORBED) s->n_absorbed += 1;
    else if (r->status == ORC_SUSPENDED) s->n_suspended += 1;
    if (port_flag_g(g, r->pos, r->status)) s->n_exit_port += 1;
}

int orc_trace(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed,
              int prec, orc_record* rec, orc_stats* stats, int n_threads) {
    geom g; consts_f kf; consts_d kd;
    if (make_geom(sc, &g, &kf, &kd)) return -1;
    double d0[3], x0[3];
    int kind0 = launch(&g, src->pos, src->dir, d0, x0);
    if (kind0 < 0) return -2;
    int nt = pick_threads(n_threads);
    orc_stats tot; memset(&tot, 0, sizeof tot);
    #pragma omp parallel num_threads(nt)
    {
        orc_stats loc; memset(&loc, 0, sizeof loc);
        #pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)n; i++) {
            orc_record r;
            run_ray(&g, &kf, &kd, prec, kind0, x0, d0, seed, ray_id0 + (uint64_t)i, NULL, &r, NULL, NULL, NULL);
            if (rec) rec[i] = r;
            add_stats(&g, &r, &loc);
        }
        #pragma omp critical
        {
            tot.n_rays += loc.n_rays; tot.n_exited += loc.n_exited; tot.n_exit_port += loc.n_exit_port;
            tot.n_absorbed += loc.n_absorbed; tot.n_suspended += loc.n_suspended; tot.n_bounces += loc.n_bounces;
        }
    }
    if (stats) *stats = tot;
    return 0;
}

int orc_trace_f64(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed,
                  double* pos, double* dir, uint32_t* n_hits, uint8_t* status, int n_threads) {
    geom g; consts_f kf; consts_d kd;
    if (make_geom(sc, &g, &kf, &kd)) return -1;
    double d0[3], x0[3];
    int kind0 = launch(&g, src->pos, src->dir, d0, x0);
    if (kind0 < 0) return -2;
    int nt = pick_threads(n_threads);
    #pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        result_d r;
        run_ray(&g, &kf, &kd, ORC_F64, kind0, x0, d0, seed, ray_id0 + (uint64_t)i, NULL, NULL, &r, NULL, NULL);
        for (int j = 0; j < 3; j++) { if (pos) pos[3 * i + j] = r.pos[j]; if (dir) dir[3 * i + j] = r.dir[j]; }
        if (n_hits) n_hits[i] = r.n_hits;
        if (status) status[i] = (uint8_t)r.status;
    }
    return 0;
}

/* mirror of altb_count_horizon (hits whose tilted normal no longer faces the incoming ray) */
int orc_count_horizon(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed, int prec,
                      uint64_t* n_events, uint64_t* n_rays_flagged, uint64_t* n_hits) {
    geom g; consts_f kf; consts_d kd;
    if (make_geom(sc, &g, &kf, &kd)) return -1;
    if (g.brdf_kind == 3) return -1;
    double d0[3], x0[3];
    int kind0 = launch(&g, src->pos, src->dir, d0, x0);
    if (kind0 < 0) return -2;
    uint64_t ev = 0, fl = 0, hh = 0;
    if (sc->roughness_rad != 0.0) {
        #pragma omp parallel for schedule(dynamic, 256) reduction(+ : ev, fl, hh)
        for (int64_t i = 0; i < (int64_t)n; i++) {
            float dr[ORC_DRAWS_PER_HIT];
            uint64_t e = 0;
            uint32_t k = 0, hits;
            if (prec == ORC_F32) {
                state_f s;
                int st = start_f(&g, &kf, &s, kind0, x0, d0);
                while (!st) {
                    if (g.brdf_kind == 2) orc_draws_lobe(seed, ray_id0 + (uint64_t)i, k, g.lobe_n, kf.lobe_ang, dr);
                    else orc_draws(seed, ray_id0 + (uint64_t)i, k, dr);
                    k++;
                    if (!(kf.rho < dr[0])) e += (uint64_t)past_horizon_f(&kf, &s, dr);
                    st = bounce_f(&g, &kf, &s, dr);
                }
                hits = s.n_hits;
            } else {
                state_d s;
                int st = start_d(&g, &kd, &s, kind0, x0, d0);
                while (!st) {
                    if (g.brdf_kind == 2) orc_draws_lobe(seed, ray_id0 + (uint64_t)i, k, g.lobe_n, kf.lobe_ang, dr);
                    else orc_draws(seed, ray_id0 + (uint64_t)i, k, dr);
                    k++;
                    if (!(kd.rho < (double)dr[0])) e += (uint64_t)past_horizon_d(&kd, &s, dr);
                    st = bounce_d(&g, &kd, &s, dr);
                }
                hits = s.n_hits;
            }
            ev += e; fl += e != 0; hh += hits;
        }
    }
    if (n_events) *n_events = ev;
    if (n_rays_flagged) *n_rays_flagged = fl;
    if (n_hits) *n_hits = hh;
    return 0;
}

int orc_replay(const orc_scene* sc, const double* ray0, const float* tape, const uint64_t* tape_off,
               uint64_t n, int prec, orc_record* rec) {
    return orc_replay_ex(sc, ray0, tape, tape_off, n, prec, 0u, rec);
}

int orc_replay_ex(const orc_scene* sc, const double* ray0, const float* tape, const uint64_t* tape_off,
                  uint64_t n, int prec, uint32_t flags, orc_record* rec) {
    geom g; consts_f kf; consts_d kd;
    if (make_geom(sc, &g, &kf, &kd)) return -1;
    g_full_az = (flags & ORC_REPLAY_FULL_AZIMUTH) != 0;     /* F32 mode only: the F64 mode always takes sin / cos of the full draw */
    int bad = 0;
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        double d0[3], x0[3];
        int kind0 = launch(&g, ray0 + 6 * i, ray0 + 6 * i + 3, d0, x0);
        if (kind0 < 0) { bad = 1; memset(&rec[i], 0, sizeof rec[i]); continue; }
        tape_src ts = {tape + 8 * tape_off[i], tape_off[i + 1] - tape_off[i]};
        run_ray(&g, &kf, &kd, prec, kind0, x0, d0, 0, 0, &ts, &rec[i], NULL, NULL, NULL);
    }
    g_full_az = 0;
    return bad ? -2 : 0;
}

int64_t orc_make_tape(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n,
                      uint64_t seed, float* tape, uint64_t cap_records, uint64_t* tape_off) {
    geom g; consts_f kf; consts_d kd;
    if (make_geom(sc, &g, &kf, &kd)) return -1;
    double d0[3], x0[3];
    int kind0 = launch(&g, src->pos, src->dir, d0, x0);
    if (kind0 < 0) return -2;
    /* pass 1: lengths */
    uint64_t total = 0;
    tape_off[0] = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t cnt = 0;
        run_ray(&g, &kf, &kd, ORC_F32, kind0, x0, d0, seed, ray_id0 + i, NULL, NULL, NULL, NULL, &cnt);
        total += cnt;
        tape_off[i + 1] = total;
    }
    if (!tape) return (int64_t)total;            /* length query: tape_off is filled */
    if (total > cap_records) return -3;
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint64_t cnt = 0;
        run_ray(&g, &kf, &kd, ORC_F32, kind0, x0, d0, seed, ray_id0 + (uint64_t)i, NULL, NULL, NULL,
                tape + 8 * tape_off[i], &cnt);
    }
    return (int64_t)total;
}

/* ------------------------------------------------------------------ map stage */
void orc_detector_pose(double theta_deg, double phi_deg, double radius, double pos[3], double nrm[3]) {
    /* fluxAtObserverFast.C:61-80, literal */
    double theta_rad = theta_deg * M_PI / 180.0;
    double phi_rad = phi_deg * M_PI / 180.0;
    double x = radius * sin(theta_rad) * cos(phi_rad);
    double y = radius * sin(theta_rad) * sin(phi_rad);
    double z = -100.0 - radius * cos(theta_rad);
    double dx = x - 0, dy = y - 0, dz = z - (-100.0);
    double mag = sqrt(dx * dx + dy * dy + dz * dz);
    pos[0] = x; pos[1] = y; pos[2] = z;
    nrm[0] = -dy / mag; nrm[1] = dx / mag; nrm[2] = dz / mag;
}

int orc_detector_hit(const double pos[3], const double nrm[3], double width,
                     const double lastPoint[3], const double direction[3]) {
    /* fluxAtObserverFast.C:82-119, literal */
    double nx = nrm[0], ny = nrm[1], nz = nrm[2];
    double dot = direction[0] * nx + direction[1] * ny + direction[2] * nz;
    if (fabs(dot) < 1e-10) return 0;
    double dx = lastPoint[0] - pos[0], dy = lastPoint[1] - pos[1], dz = lastPoint[2] - pos[2];
    double t = -(dx * nx + dy * ny + dz * nz) / dot;
    double ix = lastPoint[0] + direction[0] * t, iy = lastPoint[1] + direction[1] * t, iz = lastPoint[2] + direction[2] * t;
    double rx = ix - pos[0], ry = iy - pos[1], rz = iz - pos[2];
    double ux = ny * rz - nz * ry, uy = nz * rx - nx * rz, uz = nx * ry - ny * rx;
    double r2 = ux * ux + uy * uy + uz * uz;
    return r2 <= (width / 2) * (width / 2);
}

int32_t orc_direction_bin(const orc_map_spec* map, const float d[3]) {
    if (!(d[2] < 0.0f)) return -1;
    double c = -(double)d[2];
    if (c > 1.0) c = 1.0;
    double th = acos(c) * (180.0 / PI_D);
    double ph = atan2((double)d[1], (double)d[0]) * (180.0 / PI_D);
    if (ph < 0.0) ph += 360.0;
    int i = (int)floor(th / (90.0 / map->n_theta));
    int j = (int)floor(ph / (360.0 / map->n_phi));
    if (i > map->n_theta - 1) i = map->n_theta - 1;
    if (j > map->n_phi - 1) j = map->n_phi - 1;
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    return i * map->n_phi + j;
}

typedef struct { float* st; float* ct; float* rc2; float* cp; float* sp; float w2, RR, n2R, p2R; } map_tab_f;

static void make_tab_f(const orc_map_spec* map, map_tab_f* t) {
    int nt = map->n_theta, np = map->n_phi;
    t->st = malloc(sizeof(float) * nt); t->ct = malloc(sizeof(float) * nt); t->rc2 = malloc(sizeof(float) * nt);
    t->cp = malloc(sizeof(float) * np); t->sp = malloc(sizeof(float) * np);
    for (int i = 0; i < nt; i++) {
        double th = (i + 0.5) * 90.0 / nt * PI_D / 180.0;
        t->st[i] = (float)sin(th); t->ct[i] = (float)cos(th);
        t->rc2[i] = (float)(map->det_radius * cos(th) * cos(th));
    }
    for (int j = 0; j < np; j++) {
        double ph = (j + 0.5) * 360.0 / np * PI_D / 180.0;
        t->cp[j] = (float)cos(ph); t->sp[j] = (float)sin(ph);
    }
    double hw = map->det_width / 2;
    t->w2 = (float)(hw * hw);
    t->RR = (float)(map->det_radius * map->det_radius);
    t->n2R = (float)(-2.0 * map->det_radius); t->p2R = (float)(2.0 * map->det_radius);
}
static void free_tab_f(map_tab_f* t) { free(t->st); free(t->ct); free(t->rc2); free(t->cp); free(t->sp); }

/* The kernels' f32 line-disk test (DESIGN.md "map stage"): Detector::setPosition + checkIntersection
 * (fluxAtObserverFast.C:61-119) multiplied through by dot^2 and EXPANDED about the hemisphere centre c0 = (0, 0, -100).
 * With u = (st cp, st sp, -ct) the detector centre is c0 + R u and the reference's normal (-d_y, d_x, d_z)/|d| is
 * n = (-st sp, st cp, -ct), so that u.n = ct^2 depends on the row only and every scalar product splits into a per-column
 * and a per-row part.  The line is represented by its foot point m (relative to c0: the point of the line closest to c0,
 * which keeps |D|^2 small -- 1e4 instead of 1e5 cm^2 -- and with it the cancellation in the expanded form) and v:
 *   hit  <=>  |dot D - num v|^2 = dot^2 |D|^2 - 2 dot num (D.v) + num^2 |v|^2  <=  w^2 dot^2,   D = m - R u.
 * Against the literal double-precision formula 4e-6 of the hits differ (rim of the disk). */
typedef struct { float m[3], v[3], vv, mv2, mm; } line_f;
static inline void make_line_f32(const map_tab_f* t, const float* L, const float* v, line_f* c) {
    float Lp[3] = {L[0], L[1], L[2] + 100.0f};
    float t0 = -fmaf(Lp[0], v[0], fmaf(Lp[1], v[1], Lp[2] * v[2]));
    for (int i = 0; i < 3; i++) { c->m[i] = fmaf(t0, v[i], Lp[i]); c->v[i] = v[i]; }
    c->vv = fmaf(v[0], v[0], fmaf(v[1], v[1], v[2] * v[2]));
    c->mv2 = -2.0f * fmaf(c->m[0], v[0], fmaf(c->m[1], v[1], c->m[2] * v[2]));
    c->mm = fmaf(c->m[0], c->m[0], fmaf(c->m[1], c->m[1], fmaf(c->m[2], c->m[2], t->RR)));
}
static inline int line_hit_f32(const map_tab_f* t, int i, int j, const line_f* c) {
    float cp = t->cp[j], sp = t->sp[j], st = t->st[i], ct = t->ct[i];
    float A = fmaf(c->m[0], cp, c->m[1] * sp), B = fmaf(c->m[1], cp, -(c->m[0] * sp));
    float Cq = fmaf(c->v[0], cp, c->v[1] * sp), E = fmaf(c->v[1], cp, -(c->v[0] * sp));
    float g = c->v[2] * ct, k = c->m[2] * ct, h = k + t->rc2[i];
    float dot = fmaf(st, E, -g), num = fmaf(st, B, -h), um = fmaf(st, A, -k);
    float DD = fmaf(t->n2R, um, c->mm), uv = fmaf(st, Cq, -g), Dv2 = fmaf(t->p2R, uv, c->mv2);
    float a = dot * dot, b = dot * num, cc = num * num;
    float r2 = fmaf(a, DD, fmaf(b, Dv2, cc * c->vv));
    return fabsf(dot) >= 1e-10f && r2 <= t->w2 * a;
}

int orc_map_records(const orc_scene* sc, const orc_map_spec* map, const orc_record* rec, uint64_t n,
                    int prec, uint64_t* counts, int n_threads) {
    return orc_map_records_at(sc, map, rec, n, 0, prec, counts, n_threads);
}

int orc_map_records_at(const orc_scene* sc, const orc_map_spec* map, const orc_record* rec, uint64_t n, uint64_t ray_base,
                       int prec, uint64_t* counts, int n_threads) {
    geom g; consts_f kf; consts_d kd;
    if (make_geom(sc, &g, &kf, &kd)) return -1;
    int nt = map->n_theta, np = map->n_phi, nb = nt * np;
    if (nt < 1 || np < 1) return -1;
    if (map->map_mode == ORC_MAP_PER_POSITION || map->map_mode == ORC_MAP_TWOFOLD) {
        int two = map->map_mode == ORC_MAP_TWOFOLD, half = np / 2;
        if (map->rays_per_position < 1 || (two && (np & 1))) return -1;
        uint64_t n_groups = two ? (uint64_t)nt * half : (uint64_t)nb;
        map_tab_f tab; make_tab_f(map, &tab);
        for (uint64_t r = 0; r < n; r++) {
            if (!port_flag_g(&g, rec[r].pos, rec[r].status)) continue;
            uint64_t grp = (ray_base + r) / (uint64_t)map->rays_per_position;
            if (grp >= n_groups) continue;
            int i = two ? (int)(grp / half) : (int)(grp / np), j = two ? (int)(grp % half) : (int)(grp % np);
            for (int rep = 0; rep < (two ? 2 : 1); rep++, j += half) {
                int hit;
                if (prec == ORC_F32) { line_f lc; make_line_f32(&tab, rec[r].pos, rec[r].dir, &lc); hit = line_hit_f32(&tab, i, j, &lc); }
                else {
                    double p[3], nn[3], L[3] = {rec[r].pos[0], rec[r].pos[1], rec[r].pos[2]}, v[3] = {rec[r].dir[0], rec[r].dir[1], rec[r].dir[2]};
                    orc_detector_pose((i + 0.5) * 90.0 / nt, (j + 0.5) * 360.0 / np, map->det_radius, p, nn);
                    hit = orc_detector_hit(p, nn, map->det_width, L, v);
                }
                counts[(size_t)i * np + j] += (uint64_t)hit;
            }
        }
        free_tab_f(&tab);
        return 0;
    }
    if (map->map_mode == ORC_MAP_DIRECTION) {
        for (uint64_t r = 0; r < n; r++) {
            if (!port_flag_g(&g, rec[r].pos, rec[r].status)) continue;
            int32_t b = orc_direction_bin(map, rec[r].dir);
            if (b >= 0) counts[b] += 1;
        }
        return 0;
    }
    if (map->map_mode != ORC_MAP_LINE && map->map_mode != ORC_MAP_TRACEONCE_COMPAT) return -1;
    int compat = map->map_mode == ORC_MAP_TRACEONCE_COMPAT;
    map_tab_f tab; make_tab_f(map, &tab);
    double* dp = malloc(sizeof(double) * 6 * nb);
    for (int i = 0; i < nt; i++)
        for (int j = 0; j < np; j++)
            orc_detector_pose((i + 0.5) * 90.0 / nt, (j + 0.5) * 360.0 / np, map->det_radius,
                              dp + 6 * (i * np + j), dp + 6 * (i * np + j) + 3);
    int nthr = pick_threads(n_threads);
    #pragma omp parallel num_threads(nthr)
    {
        uint64_t* loc = calloc(nb, sizeof(uint64_t));
        #pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < (int64_t)n; r++) {
            if (!port_flag_g(&g, rec[r].pos, rec[r].status)) continue;
            if (prec == ORC_F32) {
                float L[3], v[3];
                if (compat) {
                    /* fluxAtObserverFast.C:1181 GetPoint() leaves the start at (0,0,0): line from the
                     * origin through the exit point */
                    const float* e = rec[r].pos;
                    float inv = 1.0f / sqrtf(fmaf(e[0], e[0], fmaf(e[1], e[1], e[2] * e[2])));
                    L[0] = L[1] = L[2] = 0.0f;
                    v[0] = e[0] * inv; v[1] = e[1] * inv; v[2] = e[2] * inv;
                } else {
                    for (int c = 0; c < 3; c++) { L[c] = rec[r].pos[c]; v[c] = rec[r].dir[c]; }
                }
                line_f lc;
                make_line_f32(&tab, L, v, &lc);
                for (int i = 0; i < nt; i++)
                    for (int j = 0; j < np; j++)
                        loc[i * np + j] += (uint64_t)line_hit_f32(&tab, i, j, &lc);
            } else {
                double L[3], v[3];
                if (compat) {
                    double e[3] = {rec[r].pos[0], rec[r].pos[1], rec[r].pos[2]};
                    double mag = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
                    L[0] = L[1] = L[2] = 0.0;
                    v[0] = e[0] / mag; v[1] = e[1] / mag; v[2] = e[2] / mag;
                } else {
                    for (int c = 0; c < 3; c++) { L[c] = rec[r].pos[c]; v[c] = rec[r].dir[c]; }
                }
                for (int b = 0; b < nb; b++)
                    loc[b] += (uint64_t)orc_detector_hit(dp + 6 * b, dp + 6 * b + 3, map->det_width, L, v);
            }
        }
        #pragma omp critical
        for (int b = 0; b < nb; b++) counts[b] += loc[b];
        free(loc);
    }
    free(dp); free_tab_f(&tab);
    return 0;
}

int orc_fluxmap(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed,
                const orc_map_spec* map, int prec, uint64_t* counts, orc_stats* stats, int n_threads) {
    const uint64_t chunk = 1u << 16;
    orc_record* rec = malloc(sizeof(orc_record) * chunk);
    orc_stats tot; memset(&tot, 0, sizeof tot);
    int rc = 0;
    for (uint64_t off = 0; off < n && !rc; off += chunk) {
        uint64_t m = n - off < chunk ? n - off : chunk;
        orc_stats s;
        rc = orc_trace(sc, src, ray_id0 + off, m, seed, prec, rec, &s, n_threads);
        if (!rc) rc = orc_map_records_at(sc, map, rec, m, ray_id0 + off, prec, counts, n_threads);
        tot.n_rays += s.n_rays; tot.n_exited += s.n_exited; tot.n_exit_port += s.n_exit_port;
        tot.n_absorbed += s.n_absorbed; tot.n_suspended += s.n_suspended; tot.n_bounces += s.n_bounces;
    }
    free(rec);
    if (stats) *stats = tot;
    return rc;
}

/* ------------------------------------------------------------------ physical disks */
void orc_sweep_pose(double theta, double phi, double r, double center[3], double rot[9]) {
    /* integratingSphereDetectorSweep.C:150-171; TGeoRotation::RotateZ then RotateY act in the
     * master frame: M = Ry(rotTheta) * Rz(rotPhi) */
    double x = r * sin(theta * M_PI / 180.0) * cos(phi * M_PI / 180.0);
    double y = r * sin(theta * M_PI / 180.0) * sin(phi * M_PI / 180.0);
    double z = -r * cos(theta * M_PI / 180.0);
    double dx = 0 - x, dy = 0 - y, dz = -100.0 - z;
    double rotTheta = -atan2(sqrt(dx * dx + dy * dy), dz);
    double rotPhi = atan2(dy, dx);
    double cz = cos(rotPhi), sz = sin(rotPhi), cy = cos(rotTheta), sy = sin(rotTheta);
    double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += Ry[3 * i + k] * Rz[3 * k + j];
            rot[3 * i + j] = a;
        }
    center[0] = x; center[1] = y; center[2] = z;
}

/* segment (sphere crossing -> world-box point e) against a thin cylinder; IEEE ops only */
static int disk_hit(const geom* g, const float* ef, const float* df, const double* c, const double* rot,
                    double rad, double ht) {
    double e[3] = {ef[0], ef[1], ef[2]}, d[3] = {df[0], df[1], df[2]};
    double b = (e[0] * d[0] + e[1] * d[1]) + e[2] * d[2];
    double cc = ((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]) - g->R1sq;
    double disc = b * b - cc;
    double smax = disc > 0.0 ? b - sqrt(disc) : b;
    if (!(smax > 0.0)) return 0;
    double a[3] = {rot[2], rot[5], rot[8]};
    double rel[3] = {e[0] - c[0], e[1] - c[1], e[2] - c[2]};
    double z0 = (rel[0] * a[0] + rel[1] * a[1]) + rel[2] * a[2];
    double dz = -((d[0] * a[0] + d[1] * a[1]) + d[2] * a[2]);
    double lo = 0.0, hi = smax;
    if (dz != 0.0) {
        double s0 = (-ht - z0) / dz, s1 = (ht - z0) / dz;
        if (s0 > s1) { double t = s0; s0 = s1; s1 = t; }
        if (s0 > lo) lo = s0;
        if (s1 < hi) hi = s1;
    } else if (fabs(z0) > ht) return 0;
    if (lo > hi) return 0;
    double rp[3], dp[3];
    for (int i = 0; i < 3; i++) { rp[i] = rel[i] - z0 * a[i]; dp[i] = -d[i] - dz * a[i]; }
    double qa = (dp[0] * dp[0] + dp[1] * dp[1]) + dp[2] * dp[2];
    double qb = (rp[0] * dp[0] + rp[1] * dp[1]) + rp[2] * dp[2];
    double qc = ((rp[0] * rp[0] + rp[1] * rp[1]) + rp[2] * rp[2]) - rad * rad;
    if (qa > 0.0) {
        double dd = qb * qb - qa * qc;
        if (dd < 0.0) return 0;
        double sq = sqrt(dd);
        double s0 = (-qb - sq) / qa, s1 = (-qb + sq) / qa;
        if (s0 > lo) lo = s0;
        if (s1 < hi) hi = s1;
    } else if (qc > 0.0) return 0;
    return lo <= hi;
}

int orc_disk_hits(const orc_scene* sc, const orc_record* rec, uint64_t n, const double* det_center,
                  const double* det_rot, uint32_t m, double det_r, double det_halfthick, uint64_t* hits) {
    geom g; consts_f kf; consts_d kd;
    if (make_geom(sc, &g, &kf, &kd)) return -1;
    for (uint64_t r = 0; r < n; r++) {
        if (rec[r].status != ORC_EXITED) continue;
        for (uint32_t j = 0; j < m; j++)
            hits[j] += (uint64_t)disk_hit(&g, rec[r].pos, rec[r].dir, det_center + 3 * j, det_rot + 9 * j, det_r, det_halfthick);
    }
    return 0;
}

/* ------------------------------------------------------------------ polylines (what ARay::MakePolyLine3D shows,
 * makeIntegratingSphereNRays.C:69-72): point 0 = source, then every surface hit, then the world-box point of an
 * exited ray.  F32 arithmetic (the kernels' mirror).  pts[n][max_points][3]; npts[i] is the TRUE number of points
 * (1 + hits + exited), only the first max_points are stored. */
int orc_trace_paths(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed,
                    uint32_t max_points, float* pts, uint32_t* npts, uint8_t* status) {
    geom g; consts_f kf; consts_d kd;
    if (make_geom(sc, &g, &kf, &kd)) return -1;
    if (g.brdf_kind == 3) return -1;                  /* polylines show ONE ARay: the two-ray post-hoc mode has none */
    double d0[3], x0[3];
    int kind0 = launch(&g, src->pos, src->dir, d0, x0);
    if (kind0 < 0) return -2;
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        float* p = pts + (size_t)i * max_points * 3;
        uint32_t np_ = 0;
        #define PUT(v) do { if (np_ < max_points) { p[3 * np_] = (v)[0]; p[3 * np_ + 1] = (v)[1]; p[3 * np_ + 2] = (v)[2]; } np_++; } while (0)
        float s0[3] = {(float)src->pos[0], (float)src->pos[1], (float)src->pos[2]};
        PUT(s0);
        state_f s;
        float dr[ORC_DRAWS_PER_HIT];
        uint32_t k = 0;
        int st = start_f(&g, &kf, &s, kind0, x0, d0);
        while (!st) {
            PUT(s.pos);
            if (g.brdf_kind == 2) orc_draws_lobe(seed, ray_id0 + (uint64_t)i, k, g.lobe_n, kf.lobe_ang, dr);
            else orc_draws(seed, ray_id0 + (uint64_t)i, k, dr);
            k++;
            st = bounce_f(&g, &kf, &s, dr);
        }
        if (st == ORC_EXITED) PUT(s.pos);
        #undef PUT
        npts[i] = np_;
        if (status) status[i] = (uint8_t)st;
    }
    return 0;
}
