/*
 * altair_oracle.h -- CPU restatement of the reference's integrating-sphere hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (altair-raytracing_b200/,
 * include/, macros) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker / the CPU arm.
 *
 * PARITY STATUS: "parity unpinned" at the bit level.  The arithmetic of the path lives in
 * ROOT 6.34.04 + ROBAST (un-vendored, not installable here), the
 * reference has no tests and no dumped draws.  This oracle is pinned STATISTICALLY against
 * the reference's committed outputs (tests/golden/, made by tests/golden/make_golden.py
 * from /root/reference/flux_at_observer/ CSVs, 3dRayLog.txt, angular_dist.txt).
 *
 * Two arithmetic modes of the SAME algorithm:
 *   prec = ORC_F64  straightforward double precision with libm sin/cos/log/sqrt -- the
 *                   "physics" restatement that is checked against the reference goldens;
 *   prec = ORC_F32  the documented single-precision operation sequence (DESIGN.md
 *                   "arithmetic contract": IEEE add/mul/fma/div/sqrt only, polynomial
 *                   sin/cos/log) that the CUDA kernels must reproduce bit for bit.
 * The kernels are held to BOTH: bit-exact against ORC_F32, and -- the north-star's replay
 * criterion -- per-ray status / port flag / bin equal to ORC_F64 on the same draws except for
 * <= 1e-4 of the rays (measured 1.3e-5: rounding differences do not grow along a trajectory,
 * they only flip decisions that sit within FP32 epsilon of a boundary; tests/test_oracle_modes.py).
 *
 * What it restates (reference file:line, all under /root/reference):
 *   scene            flux_at_observer/fluxAtObserverFast.C:33-41,192-230
 *                    makeIntegratingSphereNRays.C:25-39, integratingSphereDetectorSweep.C:114-123
 *   source ray       flux_at_observer/fluxAtObserverFast.C:1147-1150
 *   bounce loop      AOpticsManager::TraceNonSequential (ROBAST, call sites
 *                    fluxAtObserverFast.C:1153, fluxAtObserverOptimize.C:295,
 *                    makeIntegratingSphereNRays.C:67) -- algorithm per SURVEY.md appendix A
 *   exit criterion   flux_at_observer/fluxAtObserverOptimize.C:309,323 (lastPoint z < -100)
 *   detector         flux_at_observer/fluxAtObserverFast.C:61-80 (setPosition), :82-119
 *                    (checkIntersection)
 *   trace-once map   flux_at_observer/fluxAtObserverFast.C:1164-1303
 *   BRDF mixture     flux_at_observer/nonLambertianFlux.C:147-208
 *   post-hoc mode    flux_at_observer/nonLambertianFlux.C:235-304 (re-scatter at the last point + second trace)
 *   direction hists  distributionSphereDetectorSweep.C:74-99
 *   physical disk    integratingSphereDetectorSweep.C:134-172
 */
#ifndef ALTAIR_ORACLE_H
#define ALTAIR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order as altb_scene in include/altair_b200.h (declared independently). */
typedef struct {
    double r_inner, r_outer, theta_max_deg, world_half;
    double reflectance, roughness_rad;
    int32_t lambertian;   /* 1: diffuse model selected by brdf_kind; 0: ideal specular mirror */
    int32_t max_bounces;  /* AOpticsManager::SetLimit */
    int32_t brdf_kind;    /* 0 Lambert, 1 spec/diffuse mixture (nonLambertianFlux.C:147-208) at every bounce,
                             2 cos^n lobe ('nonLambertianFlux copy.C':31-70),
                             3 the committed macro literally (nonLambertianFlux.C:246-268): Lambertian trace, ONE sample of
                               the kind-1 mixture at the last point (normal = lastPoint.Unit(), incident = the INITIAL
                               direction), second Lambertian trace from there; the record is the second ray's, n_hits the sum */
    int32_t count_all_status; /* 0: only EXITED rays can "pass the port" (batch macros read
                                 GetExited/GetStopped only); 1: any status (single-ray macros) */
    double brdf_param[4]; /* kind 1: roughness, specular, diffuse; kind 2: exponent (integer 1..8), max angle [deg] */
    double exit_z;
} orc_scene;

typedef struct { double pos[3], dir[3]; } orc_source;

typedef struct {
    int32_t n_theta, n_phi;
    double det_radius, det_width;
    int32_t map_mode;      /* 0 LINE, 1 TRACEONCE_COMPAT, 2 DIRECTION, 3 PER_POSITION, 4 TWOFOLD */
    int32_t rays_per_position;
} orc_map_spec;

typedef struct {
    uint64_t n_rays, n_exited, n_exit_port, n_absorbed, n_suspended, n_bounces;
    double t_trace_s, t_map_s;
} orc_stats;

/* Per-ray result, identical layout to the kernels' 32-byte record. */
typedef struct { float pos[3]; float dir[3]; uint32_t n_hits; uint32_t status; } orc_record;

enum { ORC_EXITED = 1, ORC_ABSORBED = 2, ORC_SUSPENDED = 3, ORC_TAPE_END = 4 };
enum { ORC_MAP_LINE = 0, ORC_MAP_TRACEONCE_COMPAT = 1, ORC_MAP_DIRECTION = 2, ORC_MAP_PER_POSITION = 3, ORC_MAP_TWOFOLD = 4 };
enum { ORC_F64 = 0, ORC_F32 = 1 };

#define ORC_DRAWS_PER_HIT 8
/* draw record of one surface hit (f32):
 *   [0] u_abs  [1] u_r  [2] u_phi  [3] u_sel  [4] u_psi  [5] g0  [6] g1  [7] reserved (0)
 * all derived from ONE Philox4x32-10 block per hit (bit budget in altair_oracle.c:orc_draws);
 * u_* uniform on [0,1); g0,g1 independent N(0,1) (Box-Muller). */

/* Philox4x32-10 / -7 (Salmon et al. 2011), one block. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_philox4x32_7(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* rounds behind orc_draws / orc_trace / orc_fluxmap: 10 (default) or 7 (mirror of ALTB_CONTRACT_FAST7); process-wide */
int orc_set_philox_rounds(int rounds);

/* The 8 f32 draws for (seed, ray_id, hit index k) exactly as the CUDA kernels derive them. */
void orc_draws(uint64_t seed, uint64_t ray_id, uint32_t k, float out[ORC_DRAWS_PER_HIT]);
/* brdf_kind 2: slot [1] is the polar draw ACCEPTED by the cos^n rejection loop ('nonLambertianFlux copy.C':47-69) */
void orc_draws_lobe(uint64_t seed, uint64_t ray_id, uint32_t k, int lobe_n, float lobe_ang, float out[ORC_DRAWS_PER_HIT]);

/* f32 math primitives of the arithmetic contract (exposed for unit tests). */
void  orc_sincos2pi_f32(float u, float* s, float* c);
void  orc_sincos2pi_q13(uint32_t q, float* s, float* c);   /* table form, 13-bit turn fraction */
void  orc_sincos2pi_q20(uint32_t q, float* s, float* c);   /* table + second-order rotation, 20-bit turn fraction */
void  orc_sincos_f32(float x, float* s, float* c);
float orc_log_u20(uint32_t k);   /* ln(k 2^-20), k = 1 .. 2^20 */

/* Trace rays ray_id0 .. ray_id0+n-1 with Philox draws; rec and/or stats may be NULL. */
int orc_trace(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed,
              int prec, orc_record* rec, orc_stats* stats, int n_threads);
/* Same, double outputs of the ORC_F64 state (for physics checks): pos[3n], dir[3n]. */
int orc_trace_f64(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed,
                  double* pos, double* dir, uint32_t* n_hits, uint8_t* status, int n_threads);

/* SURVEY.md A.3 step 2, large roughness (fluxAtObserver.C:156): surface hits (not absorbed) at which the roughness-tilted
 * normal no longer faces the incoming ray, rays with at least one such hit, all surface hits of the rays. */
int orc_count_horizon(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed, int prec,
                      uint64_t* n_events, uint64_t* n_rays_flagged, uint64_t* n_hits);

/* Replay: ray i starts at ray0[i] = (pos, dir) and consumes tape records
 * tape[8*tape_off[i] .. 8*tape_off[i+1]).  Status ORC_TAPE_END if the tape runs out. */
int orc_replay(const orc_scene* sc, const double* ray0, const float* tape, const uint64_t* tape_off,
               uint64_t n, int prec, orc_record* rec);
/* ORC_REPLAY_FULL_AZIMUTH: the F32 mode takes sin / cos of 2 pi u_phi, 2 pi u_psi at the draws' full float precision instead
 * of truncating them to 20 / 13-bit turn fractions (mirror of altb_replay_ex; for tapes not recorded from the Philox path). */
#define ORC_REPLAY_FULL_AZIMUTH 1u
int orc_replay_ex(const orc_scene* sc, const double* ray0, const float* tape, const uint64_t* tape_off,
                  uint64_t n, int prec, uint32_t flags, orc_record* rec);

/* Generate the tape a Philox trace of the same rays would use (ORC_F32 trajectory).
 * tape == NULL: only fills tape_off and returns the number of records needed.
 * Returns number of records written; -3 when cap_records is too small. */
int64_t orc_make_tape(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n,
                      uint64_t seed, float* tape, uint64_t cap_records, uint64_t* tape_off);

/* "Escaped through the port" flag of a record (fluxAtObserverOptimize.C:309,323). */
int orc_port_flag(const orc_scene* sc, const orc_record* r);

/* Map stage: counts[n_theta*n_phi] (theta-major) += contributions of the records.
 * prec selects the arithmetic of the line-disk test (ORC_F64 = the reference's literal
 * formula in double; ORC_F32 = the kernels' division-free f32 form).  Brute force. */
int orc_map_records(const orc_scene* sc, const orc_map_spec* map, const orc_record* rec, uint64_t n,
                    int prec, uint64_t* counts, int n_threads);
/* same with the ray id of rec[0] (needed by the PER_POSITION / TWOFOLD modes: fluxAtObserverOptimize.C:542-579,
 * fluxAtObserverFast.C:660-720: ray id r is tested only against position group r / rays_per_position) */
int orc_map_records_at(const orc_scene* sc, const orc_map_spec* map, const orc_record* rec, uint64_t n, uint64_t ray_base,
                       int prec, uint64_t* counts, int n_threads);
int32_t orc_direction_bin(const orc_map_spec* map, const float d[3]);

/* trace + map in one go (chunked; multi-threaded when n_threads != 1; <= 0 -> all cores). */
int orc_fluxmap(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed,
                const orc_map_spec* map, int prec, uint64_t* counts, orc_stats* stats, int n_threads);

/* Detector pose / single test, literal double version of fluxAtObserverFast.C:61-119. */
void orc_detector_pose(double theta_deg, double phi_deg, double radius, double pos[3], double nrm[3]);
int  orc_detector_hit(const double pos[3], const double nrm[3], double width,
                      const double line_pt[3], const double line_dir[3]);

/* Physical thin-disk detectors (integratingSphereDetectorSweep.C:145-172): hits[j] += 1 when the
 * escaping ray's last segment (sphere crossing -> world box) enters disk j. */
int orc_disk_hits(const orc_scene* sc, const orc_record* rec, uint64_t n, const double* det_center,
                  const double* det_rot, uint32_t m, double det_r, double det_halfthick, uint64_t* hits);
void orc_sweep_pose(double theta_deg, double phi_deg, double r, double center[3], double rot[9]);

/* Polylines: point 0 = source, every surface hit, the world-box point of an exited ray (F32 arithmetic).
 * pts[n][max_points][3]; npts[i] = true number of points, only the first max_points are stored. */
int orc_trace_paths(const orc_scene* sc, const orc_source* src, uint64_t ray_id0, uint64_t n, uint64_t seed,
                    uint32_t max_points, float* pts, uint32_t* npts, uint8_t* status);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
