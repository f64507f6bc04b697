// altb_macros.h -- host-side mirror of the reference's ROOT macro entry points for the integrating-sphere
// hot path.  Same function names, default arguments, stdout/stderr conventions, file names and CSV text as
// the reference; the ROBAST calls (AOpticsManager::TraceNonSequential + the post-processing loops) are
// replaced by calls into the C ABI of include/altair_b200.h.  One namespace per reference macro file,
// because several files define a function of the same name.
//
//   namespace                        reference file
//   fluxAtObserverOptimize           flux_at_observer/fluxAtObserverOptimize.C   (sweepDetector :433, sweepSeries :892)
//   fluxAtObserverFast               flux_at_observer/fluxAtObserverFast.C       (sweepDetector, sweepDetectorTwofold :518,
//                                                                                 sweepDetectorTraceOnce :1068, sweepSeries :1641)
//   fluxAtObserver                   flux_at_observer/fluxAtObserver.C           (sweepDetector :231)
//   nonLambertianFlux                flux_at_observer/nonLambertianFlux.C        (sweepDetector :307)
//   makeIntegratingSphereNRays()     makeIntegratingSphereNRays.C:22
//   integratingSphereDetectorSweep() integratingSphereDetectorSweep.C:107
//   distributionSphereDetectorSweep() distributionSphereDetectorSweep.C:25
// Out of scope (interactive OpenGL): visualizeDetector, showRedRaysOnly, MakePolyLine3D drawing.
#pragma once
#include <cstdint>
#include <string>
#include "altb_th.h"

namespace altb_macros {

const double cm = 1.0;            // AOpticsManager::cm()
const double nm = 1e-7;           // AOpticsManager::nm() (wavelengths are not used by the mirror-only scene)

// File-scope constants of the reference become overridable settings (defaults = the reference's literals).
struct Settings {
    int rays_per_position = 50000;    // "int n = 50000" fluxAtObserverOptimize.C:455, fluxAtObserverFast.C:555
    int traceonce_rays = 100000;      // "int n = 100000" fluxAtObserverFast.C:1090
    int n_theta_bins = 180, n_phi_bins = 90;          // :1092-1093
    int nrays_macro = 1000;           // makeIntegratingSphereNRays.C:57
    int sweep_rays = 100000;          // integratingSphereDetectorSweep.C:125
    int distribution_rays = 10000;    // distributionSphereDetectorSweep.C:57
    int nonlambertian_rays = 100000;  // nonLambertianFlux.C:311
    int nonlambertian_posthoc = 1;    // 1: the committed macro literally -- Lambertian trace, ONE BRDF sample at the last point,
                                      //    second trace (nonLambertianFlux.C:246-268; brdf_kind 3);
                                      // 0: gBRDF at EVERY bounce ("CustomMirror", brdf_kind 1: the model of BASELINE config C3)
    double sweep_dtheta = 0.5;        // integratingSphereDetectorSweep.C:126
    uint64_t seed = 4357;             // TRandom3 default seed
    // The reference's gRandom keeps advancing from one macro call to the next, so repeated calls (the five repeats per port
    // angle of sweepSeries, fluxAtObserverFast.C:1641-1673) are independent samples.  Here a ray's random stream is
    // (seed, ray id): every macro call takes the next unused block of ray ids.  advance_ray_ids = 0 pins every call to
    // next_ray_id (reproducible single calls).
    uint64_t next_ray_id = 0;
    int advance_ray_ids = 1;
    int traceonce_as_shipped = 1;     // 1: reproduce the published trace-once maps (line from the origin);
                                      // 0: the intended semantics (true final segment)
    int contract = 0;                 // ALTB_CONTRACT_EXACT (bit-identical to the CPU oracle; default), 1 = _FAST, 2 = _FAST7
                                      // (include/altair_b200.h): the macros' maps agree statistically, not ray by ray
    int verbose = 1;
    std::string output_dir;           // prefix for relative output paths ("" = current directory, as the reference)
};
Settings& settings();

// results of the last macro call, for callers that used to keep the TH2D / counters
struct LastRun {
    std::string csv_path;
    TH2D* fluxMap = nullptr;
    TH1D* hAngularDist = nullptr; TH1D* hDirectionZ = nullptr;
    long long fluxCount = 0, totalHitRays = 0, exitedRays = 0, n_bounces = 0;
    unsigned long long first_ray_id = 0;      // ray ids [first_ray_id, first_ray_id + rays traced) of the last call
    double rayTime = 0, sweepTime = 0, totalTime = 0;
};
LastRun& last_run();

std::string getUniqueFilename(const std::string& basePath);     // fluxAtObserverOptimize.C:336-387

}  // namespace altb_macros

namespace fluxAtObserverOptimize {
void sweepDetector(bool notify = true, const char* saveFolder = "results", int threads = -1,
                   double srcX = -60 * altb_macros::cm, double srcY = 0 * altb_macros::cm, double srcZ = -80 * altb_macros::cm,
                   double dirX = 5, double dirY = 2, double dirZ = 0, double thetaMax = 170.);
void sweepSeries();
}
namespace fluxAtObserverFast {
void sweepDetector(bool notify = true, const char* saveFolder = "results", int threads = -1,
                   double srcX = -60 * altb_macros::cm, double srcY = 0 * altb_macros::cm, double srcZ = -80 * altb_macros::cm,
                   double dirX = 5, double dirY = 2, double dirZ = 0, double thetaMax = 170.);
void sweepDetectorTwofold(bool notify = true, const char* saveFolder = "results", int threads = -1,
                          double srcX = -60 * altb_macros::cm, double srcY = 0 * altb_macros::cm, double srcZ = -80 * altb_macros::cm,
                          double dirX = 5, double dirY = 2, double dirZ = 0, double thetaMax = 170.);
void sweepDetectorTraceOnce(bool notify = true, const char* saveFolder = "results", int threads = -1,
                            double srcX = -60 * altb_macros::cm, double srcY = 0 * altb_macros::cm, double srcZ = -80 * altb_macros::cm,
                            double dirX = 5, double dirY = 2, double dirZ = 0, double thetaMax = 170.);
void sweepSeries();
}
namespace fluxAtObserver { void sweepDetector(); }
namespace nonLambertianFlux { void sweepDetector(); }
void makeIntegratingSphereNRays();
void integratingSphereDetectorSweep();
void distributionSphereDetectorSweep();
