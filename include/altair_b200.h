/*
 * altair_b200.h -- C ABI of the B200-native integrating-sphere photon tracer.
 *
 * This is the drop-in boundary for ONE hot path of bdagnillo/altair-raytracing: the
 * multi-bounce trace + observer flux map that the reference's ROOT macros obtain from ROBAST
 * (AOpticsManager::TraceNonSequential) and then post-process on the host.  The reference has no
 * FFI of its own; each entry point below names the reference call sites it replaces
 * (paths relative to the reference repository root).  All structs are plain data, all
 * pointers are HOST pointers owned by the caller unless the name ends in _dev.  Every function
 * returns 0 on success or a negative ALTB_E_* code; altb_last_error() gives the message of the
 * last failure on the calling thread.  There is no CPU fallback: without a CUDA device
 * altb_create fails with ALTB_E_CUDA.
 */
#ifndef ALTAIR_B200_H
#define ALTAIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALTB_VERSION 1

enum { ALTB_OK = 0, ALTB_E_ARG = -1, ALTB_E_SCENE = -2, ALTB_E_SOURCE = -3, ALTB_E_CUDA = -4, ALTB_E_NOMEM = -5 };

/* ray status, as ARay::IsExited/IsAbsorbed/IsSuspended (fluxAtObserverFast.C:1601-1611) */
enum { ALTB_EXITED = 1, ALTB_ABSORBED = 2, ALTB_SUSPENDED = 3, ALTB_TAPE_END = 4 };

/* flux-map semantics */
enum {
    ALTB_MAP_LINE = 0,             /* Detector::checkIntersection on the true final ray,
                                      fluxAtObserverOptimize.C:82-119,302-327 */
    ALTB_MAP_TRACEONCE_COMPAT = 1, /* the map sweepDetectorTraceOnce actually produced: line from the
                                      origin through the exit point, fluxAtObserverFast.C:1164-1294 */
    ALTB_MAP_DIRECTION = 2,        /* one bin per escaping ray by exit direction (far-field limit) */
    ALTB_MAP_PER_POSITION = 3,     /* LINE test with FRESH rays per detector position: ray id r belongs to position
                                      r / rays_per_position (theta-major), fluxAtObserverOptimize.C:542-579 */
    ALTB_MAP_TWOFOLD = 4           /* as PER_POSITION, each batch shared by (theta,phi) and (theta,phi+180):
                                      group g = i*(n_phi/2)+j, fluxAtObserverFast.C:336-408,660-720 */
};

/* Scene = what setupOpticsManager builds: fluxAtObserverFast.C:33-41,192-230;
 * makeIntegratingSphereNRays.C:25-39; integratingSphereDetectorSweep.C:114-123. */
typedef struct {
    double r_inner, r_outer;      /* TGeoSphere rmin, rmax [cm] */
    double theta_max_deg;         /* TGeoSphere theta2: the port is the cone theta > theta_max */
    double world_half;            /* TGeoBBox half size [cm] */
    double reflectance;           /* AMirror::SetReflectance */
    double roughness_rad;         /* ABorderSurfaceCondition::SetGaussianRoughness */
    int32_t lambertian;           /* EnableLambertian; 0 = specular mirror */
    int32_t max_bounces;          /* AOpticsManager::SetLimit */
    int32_t brdf_kind;            /* 0 Lambert; 1 "CustomMirror" = per-bounce spec/diffuse mixture of
                                     nonLambertianFlux.C:147-208; 2 cos^n lobe of 'nonLambertianFlux copy.C':31-70;
                                     3 the committed macro literally (nonLambertianFlux.C:246-268): Lambertian trace, then ONE
                                     sample of the kind-1 mixture where the ray ended (normal = lastPoint.Unit(), incident =
                                     the INITIAL direction) and a second Lambertian trace from there -- the per-ray result is
                                     the second ray's, n_hits counts both.  Record path only (LINE-type / per-position maps,
                                     per-ray results; a DIRECTION map also goes through records); no replay, no polylines */
    int32_t count_all_status;     /* 0: only exited rays can pass the port test (batch macros);
                                     1: any final status (single-ray macros, makeIntegratingSphereNRays.C:74-78) */
    double brdf_param[4];         /* kind 1, 3: roughness, specular, diffuse (gBRDF(0.3,0.4,0.6));
                                     kind 2: exponent (integer 1..8; reference 2), max angle [deg] (reference 60) */
    double exit_z;                /* exitPortZ, -100 cm */
} altb_scene;

/* new ARay(i, lambda, x,y,z, 0, dx,dy,dz): fluxAtObserverFast.C:1147-1150 */
typedef struct { double pos[3], dir[3]; } altb_source;

/* TH2D fluxMap(n_theta,0,90; n_phi,0,360) + Detector(width,width).setPosition(theta,phi,det_radius):
 * fluxAtObserverFast.C:1092-1093,1276-1277 */
typedef struct {
    int32_t n_theta, n_phi;
    double det_radius, det_width;
    int32_t map_mode;
    int32_t rays_per_position;    /* modes PER_POSITION / TWOFOLD (int n = 50000); otherwise ignored */
} altb_map_spec;

typedef struct {
    uint64_t n_rays, n_exited, n_exit_port, n_absorbed, n_suspended, n_bounces;
    double t_trace_s, t_map_s;    /* device time of the trace / map kernels (CUDA events) */
} altb_stats;

/* Per-ray result = what the macros read back through ARay::GetLastPoint / GetDirection /
 * GetNpoints / Is*: fluxAtObserverFast.C:298-327,1164-1247. */
typedef struct { float pos[3]; float dir[3]; uint32_t n_hits; uint32_t status; } altb_record;

typedef struct altb_ctx altb_ctx;

/* devices == NULL: use device 0..n_devices-1 (n_devices <= 0: all visible devices).  A context of several devices is what
 * SetMaxThreads(k) is to the reference (fluxAtObserverFast.C:1083-1087): the rays of a call are split over the devices by
 * ray id, and the context owns one NCCL communicator per device (ncclCommInitAll; libnccl.so.2 is loaded on demand) to merge
 * the per-device maps with ONE all-reduce over NVLink -- uint64 sums, so the result is bit-identical for 1, 2, 4, 8 devices.
 * Without NCCL on the host (or with ALTB_NO_NCCL=1) the maps are summed on the host instead; altb_collective() tells which. */
int  altb_create(altb_ctx** out, const int* devices, int n_devices);
int  altb_collective(const altb_ctx* ctx);     /* 0: one device, 1: NCCL all-reduce, 2: host sum */
void altb_destroy(altb_ctx* ctx);
const char* altb_last_error(void);
int  altb_version(void);
int  altb_device_count(void);
/* batch = rays per trace launch; 0 restores the defaults: 2^28 where the trace writes 32-byte records (LINE-type maps,
 * per-ray results), 2^32 / n_scenes where it bins in the kernel (DIRECTION maps: nothing is stored per ray). */
int  altb_set_batch(altb_ctx* ctx, uint64_t batch_rays);

/* Arithmetic contract of the bounce loop.
 * ALTB_CONTRACT_EXACT (default): every FP32 result is an IEEE-754 correctly rounded operation or a table entry; the kernels
 *   equal the CPU oracle bit for bit (what the parity suite checks).
 * ALTB_CONTRACT_FAST: the same algorithm and the same random integers, with sqrt / reciprocal / log / sin / cos taken
 *   straight from the GPU's special-function unit (relative error ~1e-7).  Results agree with the exact contract the way
 *   BASELINE.json's north star defines agreement: per-ray replay against the double-precision CPU path differs for
 *   <= 1e-4 of the rays, maps agree statistically (chi^2/ndf ~ 1).  Built for brdf_kind 0 and 1 with lambertian = 1,
 *   roughness_rad <= 0.0114 (the reference's production scenes have 0.01 or 0) and a kind-1 lobe parameter <= 0.325 (0.3);
 *   other scenes, rim-aimed sources and altb_trace_paths always use the exact contract.
 * ALTB_CONTRACT_FAST7: the fast contract's arithmetic, with Philox4x32-7 instead of Philox4x32-10 as the generator (seven rounds
 *   are the fewest that pass BigCrush: Salmon et al. 2011; ten carry a safety margin).  Another random stream: results agree
 *   with the other contracts statistically, not per ray; altb_draws returns this stream and is checked bit for bit.  Replay
 *   (draws from a tape) is the fast contract itself. */
enum { ALTB_CONTRACT_EXACT = 0, ALTB_CONTRACT_FAST = 1, ALTB_CONTRACT_FAST7 = 2 };
int altb_set_contract(altb_ctx* ctx, int contract);
int altb_get_contract(const altb_ctx* ctx);

/* THE HOT PATH.  For each scene: trace rays ray_id0 .. ray_id0+n_rays-1 (ray i's random stream
 * depends only on (seed, i)) and accumulate the flux map.  counts[n_scenes][n_theta*n_phi] is
 * theta-major like the CSV rows and is ADDED to.  BATCHED SCENES: in DIRECTION mode, scenes that differ only in
 * theta_max_deg (a port-angle series, fluxAtObserverFast.C:1641-1673) share ONE persistent launch per device (up to 192
 * scenes per launch; the scene index is part of the claimed work unit), instead of one launch and one kernel tail per
 * scene; any other mix of scenes is traced group by group.  Replaces TraceNonSequential(ARayArray*) + the
 * GetExited()/checkIntersection loops: fluxAtObserverOptimize.C:281-333, fluxAtObserverFast.C:1144-1303. */
int altb_trace_fluxmap(altb_ctx* ctx, const altb_scene* scenes, int n_scenes, const altb_source* src,
                       uint64_t ray_id0, uint64_t n_rays, uint64_t seed, const altb_map_spec* map,
                       uint64_t* counts, altb_stats* stats);

/* Same on device memory of the context's first device, asynchronous on cuda_stream
 * (a cudaStream_t; NULL = default stream): d_counts[n_scenes][n_bins] uint64 is added to,
 * d_stats[n_scenes][8] uint64 is added to (n_rays, n_exited, n_exit_port, n_absorbed, n_suspended,
 * n_bounces, 0, 0).  This is what the multi-GPU driver all-reduces. */
int altb_trace_fluxmap_dev(altb_ctx* ctx, const altb_scene* scenes, int n_scenes, const altb_source* src,
                           uint64_t ray_id0, uint64_t n_rays, uint64_t seed, const altb_map_spec* map,
                           uint64_t* d_counts, uint64_t* d_stats, void* cuda_stream);

/* Per-ray results.  Replaces TraceNonSequential(ARay&) + GetLastPoint/GetDirection/GetNpoints:
 * makeIntegratingSphereNRays.C:64-78, distributionSphereDetectorSweep.C:61-99, fluxAtObserver.C:201-223.
 * Any output pointer may be NULL.  n_points = 1 + hits (+1 for the world-box point of an exited ray). */
int altb_trace_exit_rays(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0,
                         uint64_t n_rays, uint64_t seed, double* last_pos, double* last_dir,
                         uint32_t* n_points, uint8_t* status, altb_stats* stats);
int altb_trace_records(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0,
                       uint64_t n_rays, uint64_t seed, altb_record* records, altb_stats* stats);

/* Polylines for small N: what ARay::MakePolyLine3D feeds the reference's OpenGL views (makeIntegratingSphereNRays.C:69-72,
 * makeIntegratingSphere1Ray.C:21-53).  points[n][max_points][3] (f32): point 0 = source, then every surface hit, then the
 * world-box point of an exited ray; n_points[i] is the TRUE point count (= GetNpoints), only the first max_points are stored. */
int altb_trace_paths(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0, uint64_t n_rays,
                     uint64_t seed, uint32_t max_points, float* points, uint32_t* n_points, uint8_t* status);

/* Physical thin-disk detectors, traced once and tested against all m poses.  Replaces the
 * per-position re-trace of integratingSphereDetectorSweep.C:31-105,134-172.
 * det_rot[m][9] row-major TGeoRotation matrices, det_center[m][3]. hits[m] is added to.  m <= 2048 per call
 * (the reference sweeps 362 poses); split larger sweeps over calls with the same seed and ray ids. */
int altb_detector_sweep(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0,
                        uint64_t n_rays, uint64_t seed, const double* det_center, const double* det_rot,
                        uint32_t m, double det_r, double det_halfthick, uint64_t* hits, altb_stats* stats);

/* Replay mode: ray i starts at ray0[i] = (pos[3], dir[3]) and consumes the recorded draws
 * tape[8*tape_off[i] .. 8*tape_off[i+1]) (8 f32 per surface hit: u_abs,u_r,u_phi,u_sel,u_psi,g0,g1,u_spare).
 * bin[i] = direction-map bin of an escaping ray (-1 otherwise) when map != NULL. */
int altb_replay(altb_ctx* ctx, const altb_scene* scene, const double* ray0, const float* tape,
                const uint64_t* tape_off, uint64_t n_rays, const altb_map_spec* map,
                altb_record* records, int32_t* bin, uint8_t* port);
/* The azimuth draws u_phi, u_psi of a tape recorded from this library's own RNG are fixed-point turn fractions (20 and 13
 * bits) and altb_replay looks their sin / cos up, so such a tape replays bit for bit.  A tape recorded ELSEWHERE (the
 * reference's own draws, tools/root_dump_tape.C -> tools/tape_from_root_dump.py) carries arbitrary uniforms:
 * ALTB_REPLAY_FULL_AZIMUTH evaluates sin / cos of 2 pi u at the draw's full float precision instead of truncating it. */
enum { ALTB_REPLAY_FULL_AZIMUTH = 1 };
int altb_replay_ex(altb_ctx* ctx, const altb_scene* scene, const double* ray0, const float* tape,
                   const uint64_t* tape_off, uint64_t n_rays, const altb_map_spec* map, uint32_t flags,
                   altb_record* records, int32_t* bin, uint8_t* port);

/* Map stage alone on caller-provided records (host). counts is added to.  records[i] is the ray with global id ray_id0 + i:
 * the grouped modes (PER_POSITION, TWOFOLD) derive the detector position from the ray id, so a shard or batch that does not
 * start at ray 0 must say where it starts (altb_map_records assumes ray_id0 = 0). */
int altb_map_records(altb_ctx* ctx, const altb_scene* scene, const altb_map_spec* map,
                     const altb_record* records, uint64_t n, uint64_t* counts);
int altb_map_records_at(altb_ctx* ctx, const altb_scene* scene, const altb_map_spec* map,
                        const altb_record* records, uint64_t ray_id0, uint64_t n, uint64_t* counts);

/* Counter-based RNG exposed for verification: the 8 draws of hit k for rays ray_id0..+n-1. out[n][8]. */
int altb_draws(altb_ctx* ctx, uint64_t seed, uint64_t ray_id0, uint64_t n, uint32_t k, float* out);
/* same for brdf_kind 2, where slot [1] is the polar draw accepted by the cos^n rejection loop */
int altb_draws_lobe(altb_ctx* ctx, uint64_t seed, uint64_t ray_id0, uint64_t n, uint32_t k, int lobe_n, double lobe_deg, float* out);

/* f32 primitives of the arithmetic contract exposed for verification: y[i] = op(x[i]) on device 0.
 * op 0: sqrt (written-out IEEE fast path), 1: reciprocal (same), 2: natural log polynomial,
 * 3/4: sin/cos of 2 pi q / 2^20 from the azimuth table, q = (uint32) x[i]. */
int altb_probe_f32(altb_ctx* ctx, int op, const float* x, uint64_t n, float* y);

/* Diagnostic for large Gaussian roughness (fluxAtObserver.C:156, nonLambertianFlux.C:222: SetGaussianRoughness(0.5)): at how
 * many surface hits does the roughness-tilted normal no longer face the incoming ray (incoming . n_tilted >= 0)?  ROBAST does
 * not re-draw there and neither does this library (a new direction that points into the wall is mirrored about the true
 * tangent plane); SURVEY.md A.3 asks for these rays to be counted separately.  Traces the same rays with the same draws as
 * altb_trace_records: *n_events = such hits, *n_rays_flagged = rays with at least one, *n_hits = all surface hits of the
 * rays (0 when the scene has no roughness: nothing is traced).  Any output may be NULL.  Not defined for brdf_kind 3. */
int altb_count_horizon(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0,
                       uint64_t n_rays, uint64_t seed, uint64_t* n_events, uint64_t* n_rays_flagged, uint64_t* n_hits);

/* FP32 FFMA-chain throughput of device 0 [TFLOP/s] (roofline denominator measured on the box). */
int altb_measure_fp32_peak(altb_ctx* ctx, double* tflops);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t altb_launch_count(const altb_ctx* ctx);
/* ... of which launches of the bounce-loop kernel k_trace (bench.py: average launch duration of the roofline). */
uint64_t altb_trace_launch_count(const altb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
