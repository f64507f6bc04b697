// altb_macros.cpp -- the reference's macro entry points over the C ABI (see altb_macros.h).
// Text formats follow the reference byte for byte where a consumer depends on them
// (flux_at_observer/flux_analysis.py:11-57 reads '#'-comment "key: value" lines and theta,phi,fraction rows).
#include "altb_macros.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <tuple>
#include <vector>

#include "../../include/altair_b200.h"

namespace altb_macros {

Settings& settings() { static Settings s; return s; }
LastRun& last_run() { static LastRun r; return r; }

// fluxAtObserverOptimize.C:336-387 -- never overwrite: stem_1.ext, stem_2.ext, ...
std::string getUniqueFilename(const std::string& basePath) {
    FILE* file = fopen(basePath.c_str(), "r");
    if (!file) return basePath;
    fclose(file);
    std::string directory, filename;
    size_t lastSlash = basePath.find_last_of("/\\");
    if (lastSlash != std::string::npos) { directory = basePath.substr(0, lastSlash + 1); filename = basePath.substr(lastSlash + 1); }
    else { filename = basePath; }
    size_t lastDot = filename.find_last_of('.');
    std::string stem = lastDot != std::string::npos ? filename.substr(0, lastDot) : filename;
    std::string extension = lastDot != std::string::npos ? filename.substr(lastDot) : "";
    for (int counter = 1;; counter++) {
        std::string newPath = directory + stem + "_" + std::to_string(counter) + extension;
        file = fopen(newPath.c_str(), "r");
        if (!file) return newPath;
        fclose(file);
    }
}

namespace {

const int MAX_REFLECTIONS = 50000;       // fluxAtObserverFast.C:33-41
const double INNER_RADIUS = 100.1 * cm;
const double OUTER_RADIUS = 101 * cm;
const double REFLECTANCE = 0.99;
const double ROUGHNESS = 0.01;

struct Stopwatch {                        // TStopwatch::RealTime
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double RealTime() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

std::string nowString(const char* fmt = "%Y-%m-%d %H:%M:%S") {
    time_t now = time(nullptr);
    char buf[80];
    strftime(buf, sizeof buf, fmt, localtime(&now));
    return buf;
}

std::string outPath(const std::string& rel) {
    const std::string& d = settings().output_dir;
    if (d.empty() || (!rel.empty() && rel[0] == '/')) return rel;
    return d + "/" + rel;
}

// setupOpticsManager(manager, MAX_REFLECTIONS, ROUGHNESS, REFLECTANCE, thetaMax): fluxAtObserverFast.C:192-230
altb_scene fastScene(double thetaMax) {
    altb_scene s;
    memset(&s, 0, sizeof s);
    s.r_inner = INNER_RADIUS; s.r_outer = OUTER_RADIUS; s.theta_max_deg = thetaMax; s.world_half = 300 * cm;
    s.reflectance = REFLECTANCE; s.roughness_rad = ROUGHNESS; s.lambertian = 1; s.max_bounces = MAX_REFLECTIONS;
    s.brdf_kind = 0; s.count_all_status = 0; s.exit_z = -100 * cm;
    return s;
}

// the scene of the top-level macros: box 200, default reflectance (1), no roughness, limit 10000
// (makeIntegratingSphereNRays.C:25-39); single-ray macros test every ray's last point (:74-78)
altb_scene simpleScene(double thetaMax, double rOuter, double roughness) {
    altb_scene s = fastScene(thetaMax);
    s.r_outer = rOuter; s.world_half = 200 * cm; s.reflectance = 1.0; s.roughness_rad = roughness;
    s.max_bounces = 10000; s.count_all_status = 1;
    return s;
}

// the block of ray ids of this macro call (Settings::next_ray_id)
uint64_t takeRays(uint64_t n) {
    Settings& S = settings();
    const uint64_t first = S.next_ray_id;
    if (S.advance_ray_ids) S.next_ray_id += n;
    last_run().first_ray_id = first;
    return first;
}

altb_source makeSource(double x, double y, double z, double dx, double dy, double dz) {
    altb_source s = {{x, y, z}, {dx, dy, dz}};
    return s;
}

// One library context per device count, kept for the life of the process: the reference builds and deletes an AOpticsManager
// per macro call, but a context owns device buffers (records, line lists) and, with several GPUs, NCCL communicators --
// sweepSeries (25 calls) would otherwise allocate and free them 25 times.  Single host thread, as the reference.
struct CtxCache {
    altb_ctx* h[65] = {nullptr};
    void release() { for (auto& c : h) if (c) { altb_destroy(c); c = nullptr; } }
    // no destructor on purpose: at static-destruction time the CUDA runtime may already be gone; altbm_release() frees explicitly
};
CtxCache& ctx_cache() { static CtxCache c; return c; }

struct Ctx {
    altb_ctx* h = nullptr;
    explicit Ctx(int threads) {
        // the reference's `threads` argument (SetMaxThreads) selects how many GPUs of this process work on the rays
        int n = altb_device_count();
        if (threads > 0 && threads < n) n = threads;
        if (n < 1) n = 1;
        if (n > 64) n = 64;
        altb_ctx*& slot = ctx_cache().h[n];
        if (!slot) {
            if (altb_create(&slot, nullptr, n) != 0) {
                std::cerr << "Error: " << altb_last_error() << std::endl;
                slot = nullptr;
            }
            // 2^26-ray launches (2 GiB of records) are plenty for the macros' ray counts; the library default (2^28) is for long runs
            else altb_set_batch(slot, 1ull << 26);
        }
        h = slot;
        if (h && altb_set_contract(h, settings().contract) != 0) {      // Settings::contract: the arithmetic contract of every call
            std::cerr << "Error: " << altb_last_error() << std::endl;
            h = nullptr;
        }
    }
};

void fillFluxMap(TH2D* h, const std::vector<uint64_t>& counts, int nTheta, int nPhi, double n) {
    for (int i = 0; i < nTheta; i++)
        for (int j = 0; j < nPhi; j++) h->SetBinContent(i + 1, j + 1, double(counts[(size_t)i * nPhi + j]) / n);
}

TH2D* newFluxMap(const char* title, int nTheta, int nPhi) {
    LastRun& lr = last_run();
    delete lr.fluxMap;
    lr.fluxMap = new TH2D("fluxMap", title, nTheta, 0, 90, nPhi, 0, 360);
    return lr.fluxMap;
}

void mkdirP(const char* saveFolder) {
    std::string mkdirCmd = "mkdir -p \"" + outPath(saveFolder) + "\"";
    int result = system(mkdirCmd.c_str());
    if (result != 0) std::cerr << "Warning: Could not create directory: " << saveFolder << std::endl;
    else if (settings().verbose) std::cout << "Using directory: " << saveFolder << std::endl;
}

// common header block of the Fast/Optimize CSV files (fluxAtObserverOptimize.C:504-518)
void writeHeader(std::ostream& os, const char* first, const std::string& stamp, const char* nline, int n, double thetaMax, int nTheta, int nPhi,
                 double srcX, double srcY, double srcZ, double dirX, double dirY, double dirZ, const char* method) {
    os << first << stamp << std::endl;
    os << nline << n << std::endl;
    os << "# Detector dimensions: " << 40 * cm / cm << "cm x " << 40 * cm / cm << "cm" << std::endl;
    os << "# Sphere inner radius: " << INNER_RADIUS / cm << "cm" << std::endl;
    os << "# Sphere outer radius: " << OUTER_RADIUS / cm << "cm" << std::endl;
    os << "# Exit port angle: " << thetaMax << " degrees" << std::endl;
    os << "# Theta bins: " << nTheta << std::endl;
    os << "# Phi bins: " << nPhi << std::endl;
    os << "# Mirror reflectance: " << REFLECTANCE << std::endl;
    os << "# Gaussian roughness: " << ROUGHNESS << std::endl;
    os << "# Lambertian scattering: enabled" << std::endl;
    os << "# Source position (x,y,z): " << srcX / cm << "cm, " << srcY / cm << "cm, " << srcZ / cm << "cm" << std::endl;
    os << "# Source direction (x,y,z): " << dirX << ", " << dirY << ", " << dirZ << std::endl;
    os << "# Max reflections: " << MAX_REFLECTIONS << std::endl;
    if (method) os << method << std::endl;
    os << "theta,phi,fraction" << std::endl;
}

std::string mapFileName(const char* prefix, int n, int nTheta, int nPhi, double srcX, double srcY, double srcZ) {
    return std::string(prefix) + std::to_string(n) + "rays_" + std::to_string(nTheta) + "x" + std::to_string(nPhi) + "_src" +
           std::to_string(int(srcX / cm)) + "_" + std::to_string(int(srcY / cm)) + "_" + std::to_string(int(srcZ / cm)) + ".csv";
}

// CSV of sweepDetectorTraceOnce: header :1113-1128, rows :1318-1340, footer (a NEW stream, so default float
// formatting) :1374-1382
bool writeTraceOnceFile(const std::string& fullPath, const std::string& startStamp, const uint64_t* counts, int n, int nThetaBins,
                        int nPhiBins, double thetaMax, double srcX, double srcY, double srcZ, double dirX, double dirY, double dirZ,
                        long long nExit, double rayTime, double sweepTime, double totalTime) {
    std::stringstream csvBuffer;
    writeHeader(csvBuffer, "# Flux Map Data (Trace-Once Method) - Generated: ", startStamp, "# Number of rays: ", n, thetaMax, nThetaBins,
                nPhiBins, srcX, srcY, srcZ, dirX, dirY, dirZ, "# Method: Trace-Once (single trace, multiple detector positions)");
    for (int i = 0; i < nThetaBins; i++) {
        double theta = (i + 0.5) * 90.0 / nThetaBins;
        for (int j = 0; j < nPhiBins; j++) {
            double phi = (j + 0.5) * 360.0 / nPhiBins;
            double fraction = double(counts[(size_t)i * nPhiBins + j]) / double(n);
            csvBuffer << std::fixed << std::setprecision(6) << theta << "," << phi << "," << fraction << std::endl;
        }
    }
    {
        std::ofstream csvFile(fullPath, std::ios::trunc);
        if (!csvFile.is_open()) { std::cerr << "Error: Could not open file " << fullPath << " for writing." << std::endl; return false; }
        csvFile << csvBuffer.str();
        csvFile.close();
    }
    std::ofstream csvFileAppend(fullPath, std::ios::app);
    if (csvFileAppend.is_open()) {
        csvFileAppend << "# Sweep completed at: " << nowString() << std::endl;
        csvFileAppend << "# Total execution time: " << totalTime << " seconds" << std::endl;
        csvFileAppend << "# Ray tracing time: " << rayTime << " seconds" << std::endl;
        csvFileAppend << "# Detector sweep time: " << sweepTime << " seconds" << std::endl;
        csvFileAppend << "# Total rays exiting port: " << nExit << " out of " << n << std::endl;
        csvFileAppend.close();
    }
    return true;
}

// CSV of the per-position sweeps: one stream for everything, so the footer inherits std::fixed (:575-579,:667-670)
bool writePerPositionFile(const std::string& fullPath, const std::string& startStamp, bool twofold, const uint64_t* counts, int n,
                          int nThetaBins, int nPhiBins, double thetaMax, double srcX, double srcY, double srcZ, double dirX, double dirY,
                          double dirZ, double realTime, long long* totalHitsOut) {
    std::ofstream csvFile(fullPath);
    if (!csvFile.is_open()) { std::cerr << "Error: Could not open file " << fullPath << " for writing." << std::endl; return false; }
    writeHeader(csvFile, twofold ? "# Flux Map Data (Twofold Method) - Generated: " : "# Flux Map Data - Generated: ", startStamp,
                "# Number of rays per position: ", n, thetaMax, nThetaBins, nPhiBins, srcX, srcY, srcZ, dirX, dirY, dirZ,
                twofold ? "# Method: Twofold (two detectors 180\xC2\xB0 apart)" : nullptr);
    const int totalPositions = nThetaBins * nPhiBins;
    long long totalHitRays = 0;
    for (int i = 0; i < nThetaBins; i++) {
        double theta = (i + 0.5) * 90.0 / nThetaBins;
        for (int j = 0; j < (twofold ? nPhiBins / 2 : nPhiBins); j++) {
            double phi1 = (j + 0.5) * 360.0 / nPhiBins;
            double fraction1 = double(counts[(size_t)i * nPhiBins + j]) / double(n);
            totalHitRays += (long long)counts[(size_t)i * nPhiBins + j];
            csvFile << std::fixed << std::setprecision(6) << theta << "," << phi1 << "," << fraction1 << '\n';   // same bytes as std::endl, no write() per row
            if (twofold) {
                int j2 = j + nPhiBins / 2;
                double phi2 = phi1 + 180.0;
                if (phi2 >= 360.0) phi2 -= 360.0;
                double fraction2 = double(counts[(size_t)i * nPhiBins + j2]) / double(n);
                totalHitRays += (long long)counts[(size_t)i * nPhiBins + j2];
                csvFile << std::fixed << std::setprecision(6) << theta << "," << phi2 << "," << fraction2 << '\n';
            }
        }
    }
    csvFile << "# Sweep completed at: " << nowString() << std::endl;
    csvFile << "# Total execution time: " << realTime << " seconds" << std::endl;
    csvFile << "# Total ray hits: " << totalHitRays << " out of " << ((long long)n * totalPositions) << std::endl;
    csvFile.close();
    if (totalHitsOut) *totalHitsOut = totalHitRays;
    return true;
}

// per-position sweep shared by fluxAtObserverOptimize::sweepDetector, fluxAtObserverFast::sweepDetector (twofold = false)
// and fluxAtObserverFast::sweepDetectorTwofold (twofold = true)
void perPositionSweep(bool twofold, bool notify, const char* saveFolder, int threads, double srcX, double srcY, double srcZ,
                      double dirX, double dirY, double dirZ, double thetaMax) {
    Settings& S = settings();
    LastRun& lr = last_run();
    Ctx ctx(threads);
    if (!ctx.h) return;
    const int n = S.rays_per_position, nThetaBins = S.n_theta_bins, nPhiBins = S.n_phi_bins;
    mkdirP(saveFolder);
    std::string fullPath = outPath(std::string(saveFolder) + "/" +
                                   mapFileName(twofold ? "fluxmap_twofold_" : "fluxmap_", n, nThetaBins, nPhiBins, srcX, srcY, srcZ));
    fullPath = getUniqueFilename(fullPath);
    {   // fail early like the reference (:484-488) if the file cannot be created
        std::ofstream probe(fullPath);
        if (!probe.is_open()) {
            std::cerr << "Error: Could not open file " << fullPath << " for writing." << std::endl;
            return;
        }
    }
    const std::string startStamp = nowString();
    TH2D* fluxMap = newFluxMap("Detector Flux Map;#theta (deg);#phi (deg)", nThetaBins, nPhiBins);
    const int totalPositions = nThetaBins * nPhiBins;
    const long long runs = twofold ? (long long)nThetaBins * (nPhiBins / 2) : (long long)totalPositions;
    if (S.verbose)
        std::cout << "\nStarting " << (twofold ? "twofold " : "") << "detector sweep with " << n << " rays per position "
                  << "(" << totalPositions << " positions total)..." << std::endl;
    Stopwatch timer;
    altb_scene scene = fastScene(thetaMax);
    altb_source src = makeSource(srcX, srcY, srcZ, dirX, dirY, dirZ);
    altb_map_spec map = {nThetaBins, nPhiBins, 100 * cm, 40 * cm, twofold ? ALTB_MAP_TWOFOLD : ALTB_MAP_PER_POSITION, n};
    std::vector<uint64_t> counts((size_t)totalPositions, 0);
    altb_stats st;
    if (altb_trace_fluxmap(ctx.h, &scene, 1, &src, takeRays((uint64_t)runs * (uint64_t)n), (uint64_t)runs * (uint64_t)n, S.seed, &map, counts.data(), &st) != 0) {
        std::cerr << "Error: " << altb_last_error() << std::endl;
        return;
    }
    fillFluxMap(fluxMap, counts, nThetaBins, nPhiBins, double(n));
    double realTime = timer.RealTime();
    long long totalHitRays = 0;
    if (!writePerPositionFile(fullPath, startStamp, twofold, counts.data(), n, nThetaBins, nPhiBins, thetaMax, srcX, srcY, srcZ, dirX, dirY,
                              dirZ, realTime, &totalHitRays)) return;
    lr.csv_path = fullPath; lr.totalHitRays = totalHitRays; lr.totalTime = realTime; lr.rayTime = st.t_trace_s; lr.sweepTime = st.t_map_s;
    lr.exitedRays = (long long)st.n_exit_port; lr.n_bounces = (long long)st.n_bounces;
    std::cout << "\nFlux map data saved to '" << fullPath << "'" << std::endl;
    std::cout << "Sweep completed in " << realTime << " seconds (wall clock)" << std::endl;
    if (twofold)
        std::cout << "Efficiency gain: ~2x (processed " << totalPositions << " positions with " << runs << " simulation runs)" << std::endl;
    if (notify) {
        std::cout << "\n***** SWEEP COMPLETE *****\n" << std::endl;
        std::cout << '\a' << std::endl;
    }
}

}  // namespace
}  // namespace altb_macros

using namespace altb_macros;

// ------------------------------------------------------------------ fluxAtObserverOptimize.C
void fluxAtObserverOptimize::sweepDetector(bool notify, const char* saveFolder, int threads, double srcX, double srcY,
                                           double srcZ, double dirX, double dirY, double dirZ, double thetaMax) {
    perPositionSweep(false, notify, saveFolder, threads, srcX, srcY, srcZ, dirX, dirY, dirZ, thetaMax);
}

void fluxAtObserverOptimize::sweepSeries() {     // :892-921
    const double srcX = -60 * cm, srcY = 0 * cm, srcZ = -75 * cm, dirXBase = 5;
    std::string baseFolder = "results_overnight_04_1" + std::to_string(int(srcX / cm)) + "_" + std::to_string(int(srcY / cm)) + "_" +
                             std::to_string(int(srcZ / cm)) + "_" + std::to_string(int(dirXBase));
    for (double portAngle : {163., 166., 169., 172., 175., 178.})
        sweepDetector(false, baseFolder.c_str(), 1, srcX, srcY, srcZ, dirXBase, 0, 0, portAngle);
}

// ------------------------------------------------------------------ fluxAtObserverFast.C
void fluxAtObserverFast::sweepDetector(bool notify, const char* saveFolder, int threads, double srcX, double srcY, double srcZ,
                                       double dirX, double dirY, double dirZ, double thetaMax) {
    perPositionSweep(false, notify, saveFolder, threads, srcX, srcY, srcZ, dirX, dirY, dirZ, thetaMax);
}

void fluxAtObserverFast::sweepDetectorTwofold(bool notify, const char* saveFolder, int threads, double srcX, double srcY,
                                              double srcZ, double dirX, double dirY, double dirZ, double thetaMax) {
    perPositionSweep(true, notify, saveFolder, threads, srcX, srcY, srcZ, dirX, dirY, dirZ, thetaMax);
}

// :1068-1397
void fluxAtObserverFast::sweepDetectorTraceOnce(bool notify, const char* saveFolder, int threads, double srcX, double srcY,
                                                double srcZ, double dirX, double dirY, double dirZ, double thetaMax) {
    Settings& S = settings();
    LastRun& lr = last_run();
    Stopwatch setupTimer;
    Ctx ctx(threads);
    if (!ctx.h) return;
    const int n = S.traceonce_rays, nThetaBins = S.n_theta_bins, nPhiBins = S.n_phi_bins;
    mkdirP(saveFolder);
    std::string fullPath = outPath(std::string(saveFolder) + "/" + mapFileName("fluxmap_traceonce_", n, nThetaBins, nPhiBins, srcX, srcY, srcZ));
    fullPath = getUniqueFilename(fullPath);
    const std::string startStamp = nowString();
    TH2D* fluxMap = newFluxMap("Detector Flux Map (Trace-Once Method);#theta (deg);#phi (deg)", nThetaBins, nPhiBins);
    altb_scene scene = fastScene(thetaMax);
    altb_source src = makeSource(srcX, srcY, srcZ, dirX, dirY, dirZ);
    altb_map_spec map = {nThetaBins, nPhiBins, 100 * cm, 40 * cm, S.traceonce_as_shipped ? ALTB_MAP_TRACEONCE_COMPAT : ALTB_MAP_LINE, 0};
    std::vector<uint64_t> counts((size_t)nThetaBins * nPhiBins, 0);
    altb_stats st;
    if (altb_trace_fluxmap(ctx.h, &scene, 1, &src, takeRays((uint64_t)n), (uint64_t)n, S.seed, &map, counts.data(), &st) != 0) {
        std::cerr << "Error: " << altb_last_error() << std::endl;
        return;
    }
    const double rayTime = st.t_trace_s, sweepTime = st.t_map_s;
    std::cout << "Total rays exiting port: " << st.n_exit_port << " out of " << n << std::endl;
    fillFluxMap(fluxMap, counts, nThetaBins, nPhiBins, double(n));
    // (the reference adds setupTimer.RealTime() of a still-running timer here, which double-counts: :1343-1346)
    const double totalTime = setupTimer.RealTime();
    if (!writeTraceOnceFile(fullPath, startStamp, counts.data(), n, nThetaBins, nPhiBins, thetaMax, srcX, srcY, srcZ, dirX, dirY, dirZ,
                            (long long)st.n_exit_port, rayTime, sweepTime, totalTime)) return;
    lr.csv_path = fullPath; lr.exitedRays = (long long)st.n_exit_port; lr.rayTime = rayTime; lr.sweepTime = sweepTime; lr.totalTime = totalTime;
    lr.n_bounces = (long long)st.n_bounces; lr.totalHitRays = 0;
    for (uint64_t c : counts) lr.totalHitRays += (long long)c;
    std::cout << "\nFlux map data saved to '" << fullPath << "'" << std::endl;
    std::cout << "Ray tracing completed in " << rayTime << " seconds" << std::endl;
    std::cout << "Detector sweep completed in " << sweepTime << " seconds" << std::endl;
    std::cout << "Total execution time: " << totalTime << " seconds" << std::endl;
    if (notify) {
        std::cout << "\n***** TRACE-ONCE SWEEP COMPLETE *****\n" << std::endl;
        std::cout << '\a' << std::endl;
    }
}

void fluxAtObserverFast::sweepSeries() {        // :1641-1673
    const double srcX = -60 * cm, srcY = 0 * cm, srcZ = -75 * cm, dirXBase = 5, portAngle = 164.0;
    std::string baseFolder = "portAngleSweep_04_03_" + std::to_string(int(srcX / cm)) + "_" + std::to_string(int(srcY / cm)) + "_" +
                             std::to_string(int(srcZ / cm)) + "_" + std::to_string(int(portAngle));
    int n = 5;
    for (int i = 0; i < n; i++)
        sweepDetectorTraceOnce(false, baseFolder.c_str(), 1, srcX, srcY, srcZ, dirXBase, 0, 0, portAngle);
    std::cout << "\n***** ALL SWEEP SERIES COMPLETE *****\n" << std::endl;
    std::cout << '\a' << std::endl;
}

// ------------------------------------------------------------------ fluxAtObserver.C:231-406 / nonLambertianFlux.C:307-387
namespace {
// both: box 200, default reflectance, roughness 0.5, limit 10000, default Detector() = 10 cm, src (-60,0,-80), fresh rays per position
void legacySweep(bool nonLambertian) {
    Settings& S = settings();
    LastRun& lr = last_run();
    Ctx ctx(-1);
    if (!ctx.h) return;
    const int n = nonLambertian ? S.nonlambertian_rays : S.rays_per_position;
    const int nThetaBins = nonLambertian ? 45 : S.n_theta_bins, nPhiBins = nonLambertian ? 20 : S.n_phi_bins;
    const double thetaMax = 170.;
    std::string fullPath;
    std::ofstream csvFile;
    std::string timeBuffer = nowString();
    if (!nonLambertian) {
        const char* saveFolder = "results";
        mkdirP(saveFolder);
        fullPath = getUniqueFilename(outPath(std::string(saveFolder) + "/fluxmap_data_" + std::to_string(n) + "rays_" +
                                             std::to_string(nThetaBins * nPhiBins) + "points.csv"));
        csvFile.open(fullPath);
        if (!csvFile.is_open()) { std::cerr << "Error: Could not open file " << fullPath << " for writing." << std::endl; return; }
        csvFile << "# Flux Map Data - Generated: " << timeBuffer << std::endl;
        csvFile << "# Number of rays per position: " << n << std::endl;
        csvFile << "# Detector dimensions: 10cm x 10cm" << std::endl;
        csvFile << "# Sphere inner radius: 100.1cm" << std::endl;
        csvFile << "# Sphere outer radius: 101cm" << std::endl;
        csvFile << "# Exit port angle: " << thetaMax << " degrees" << std::endl;
        csvFile << "# Theta bins: " << nThetaBins << std::endl;
        csvFile << "# Phi bins: " << nPhiBins << std::endl;
        csvFile << "# y direction: 2" << std::endl;
        csvFile << "theta,phi,fraction" << std::endl;
    }
    TH2D* fluxMap = newFluxMap("Detector Flux Map;#theta (deg);#phi (deg)", nThetaBins, nPhiBins);
    altb_scene scene = simpleScene(thetaMax, 101 * cm, 0.5);
    if (nonLambertian) {      // gBRDF(0.3, 0.4, 0.6) (nonLambertianFlux.C:211): once, after the trace, as the committed macro does
        scene.brdf_kind = S.nonlambertian_posthoc ? 3 : 1;      // (:246-268) -- or at every bounce ("CustomMirror")
        scene.brdf_param[0] = 0.3; scene.brdf_param[1] = 0.4; scene.brdf_param[2] = 0.6;
    }
    altb_source src = nonLambertian ? makeSource(-60 * cm, 0, -80 * cm, 5, 0, 0) : makeSource(-60 * cm, 0, -80 * cm, 5, 2, 0);
    altb_map_spec map = {nThetaBins, nPhiBins, 100 * cm, 10 * cm, ALTB_MAP_PER_POSITION, n};
    std::vector<uint64_t> counts((size_t)nThetaBins * nPhiBins, 0);
    altb_stats st;
    if (S.verbose) {
        std::cout << "\nStarting detector sweep..." << std::endl;
        std::cout << "Format: theta(\xC2\xB0), phi(\xC2\xB0): hits/total = fraction" << std::endl;
        std::cout << "----------------------------------------" << std::endl;
    }
    if (altb_trace_fluxmap(ctx.h, &scene, 1, &src, takeRays((uint64_t)n * nThetaBins * nPhiBins), (uint64_t)n * nThetaBins * nPhiBins, S.seed, &map, counts.data(), &st) != 0) {
        std::cerr << "Error: " << altb_last_error() << std::endl;
        return;
    }
    fillFluxMap(fluxMap, counts, nThetaBins, nPhiBins, double(n));
    if (nonLambertian) {      // nonLambertianFlux.C:371-384
        fullPath = outPath("fluxmap_data.csv");
        csvFile.open(fullPath);
        csvFile << "theta,phi,fraction\n";
    }
    lr.totalHitRays = 0;
    for (int i = 0; i < nThetaBins; i++) {
        double theta = (i + 0.5) * 90.0 / nThetaBins;
        for (int j = 0; j < nPhiBins; j++) {
            double phi = (j + 0.5) * 360.0 / nPhiBins;
            double fraction = fluxMap->GetBinContent(i + 1, j + 1);
            lr.totalHitRays += (long long)counts[(size_t)i * nPhiBins + j];
            csvFile << std::fixed << std::setprecision(6) << theta << "," << phi << "," << fraction << (nonLambertian ? "\n" : "");
            if (!nonLambertian) csvFile << std::endl;
        }
    }
    if (!nonLambertian) csvFile << "# Sweep completed at: " << timeBuffer << std::endl;
    csvFile.close();
    lr.csv_path = fullPath; lr.exitedRays = (long long)st.n_exit_port; lr.n_bounces = (long long)st.n_bounces;
    lr.rayTime = st.t_trace_s; lr.sweepTime = st.t_map_s;
    std::cout << "\nFlux map data saved to '" << fullPath << "'" << std::endl;
}
}  // namespace

void fluxAtObserver::sweepDetector() { legacySweep(false); }
void nonLambertianFlux::sweepDetector() { legacySweep(true); }

// ------------------------------------------------------------------ makeIntegratingSphereNRays.C:22-100
void makeIntegratingSphereNRays() {
    Ctx ctx(-1);
    if (!ctx.h) return;
    int n = settings().nrays_macro;
    altb_scene scene = simpleScene(170., 101 * cm, 0.0);
    altb_source src = makeSource(-60 * cm, 0 * cm, -80 * cm, 5, 0, 0);
    altb_stats st;
    if (altb_trace_exit_rays(ctx.h, &scene, &src, takeRays((uint64_t)n), (uint64_t)n, settings().seed, nullptr, nullptr, nullptr, nullptr, &st) != 0) {
        std::cerr << "Error: " << altb_last_error() << std::endl;
        return;
    }
    last_run().fluxCount = (long long)st.n_exit_port;      // lastPoint[2] < exitPortZ (:74-78)
    last_run().n_bounces = (long long)st.n_bounces;
    std::cout << "Flux of rays through the exit port: " << st.n_exit_port << std::endl;
}

// ------------------------------------------------------------------ integratingSphereDetectorSweep.C:31-172
namespace {
void sweepPose(double theta, double phi, double r, double* c, double* rot) {     // addDetectorDisk :145-172
    double x = r * std::sin(theta * M_PI / 180.0) * std::cos(phi * M_PI / 180.0);
    double y = r * std::sin(theta * M_PI / 180.0) * std::sin(phi * M_PI / 180.0);
    double z = -r * std::cos(theta * M_PI / 180.0);
    double dx = 0 - x, dy = 0 - y, dz = -100 * cm - z;
    double rotTheta = -std::atan2(std::sqrt(dx * dx + dy * dy), dz), rotPhi = std::atan2(dy, dx);   // radians here
    // TGeoRotation::RotateZ(rotPhi) then RotateY(rotTheta), both in the master frame: M = Ry * Rz
    double cz = std::cos(rotPhi), sz = std::sin(rotPhi), cy = std::cos(rotTheta), sy = std::sin(rotTheta);
    double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += Ry[3 * i + k] * Rz[3 * k + j];
            rot[3 * i + j] = a;
        }
    c[0] = x; c[1] = y; c[2] = z;
}
}  // namespace

void integratingSphereDetectorSweep() {
    Ctx ctx(-1);
    if (!ctx.h) return;
    Settings& S = settings();
    const double tmax = 170.;
    const int nRays = S.sweep_rays;
    const double dtheta = S.sweep_dtheta, thetaMax = 45, dphi = 180, diskRadius = 5 * cm;
    altb_scene scene = simpleScene(tmax, 105 * cm, 0.0);       // shell 100.1 -> 105 (:119)
    altb_source src = makeSource(-60 * cm, 0, -80 * cm, 5, 0, 0);
    std::vector<double> centers, rots, thetas, phis;
    for (double theta = -thetaMax; theta <= thetaMax; theta += dtheta)
        for (double phi = 0; phi < 360; phi += dphi) {
            double c[3], m[9];
            sweepPose(theta, phi, 200 * cm, c, m);
            centers.insert(centers.end(), c, c + 3); rots.insert(rots.end(), m, m + 9);
            thetas.push_back(theta); phis.push_back(phi);
        }
    const uint32_t m = (uint32_t)thetas.size();
    std::vector<uint64_t> hits(m, 0);
    altb_stats st;
    // the reference re-traces nRays for every position; here one trace serves all positions
    if (altb_detector_sweep(ctx.h, &scene, &src, takeRays((uint64_t)nRays), (uint64_t)nRays, S.seed, centers.data(), rots.data(), m, diskRadius, 0.1 * cm,
                            hits.data(), &st) != 0) {
        std::cerr << "Error: " << altb_last_error() << std::endl;
        return;
    }
    std::string path = outPath("detector_sweep3.txt");
    std::ofstream outFile(path);
    outFile << "Theta(deg)\tPhi(deg)\tHitFraction\n";
    LastRun& lr = last_run();
    delete lr.fluxMap;
    lr.fluxMap = new TH2D("hSweepMap", "Hit Fraction Map;Theta (deg);Phi (deg)", int(2 * thetaMax / dtheta), -thetaMax, thetaMax, int(360 / dphi), 0, 360);
    for (uint32_t k = 0; k < m; k++) {
        double hitFraction = static_cast<double>(hits[k]) / nRays;
        if (S.verbose) std::cout << "Theta: " << thetas[k] << "\xC2\xB0 Phi: " << phis[k] << "\xC2\xB0 Hit fraction: " << hitFraction << std::endl;
        outFile << thetas[k] << "\t" << phis[k] << "\t" << hitFraction << "\n";
        lr.fluxMap->Fill(thetas[k], phis[k], hitFraction);
    }
    outFile.close();
    lr.csv_path = path; lr.exitedRays = (long long)st.n_exit_port; lr.n_bounces = (long long)st.n_bounces;
}

// ------------------------------------------------------------------ distributionSphereDetectorSweep.C:25-130
void distributionSphereDetectorSweep() {
    Ctx ctx(-1);
    if (!ctx.h) return;
    LastRun& lr = last_run();
    int n = settings().distribution_rays;
    altb_scene scene = simpleScene(170., 101 * cm, 0.0);
    altb_source src = makeSource(-60 * cm, 0 * cm, -80 * cm, 5, 0, 0);
    std::vector<double> pos((size_t)n * 3), dir((size_t)n * 3);
    if (altb_trace_exit_rays(ctx.h, &scene, &src, takeRays((uint64_t)n), (uint64_t)n, settings().seed, pos.data(), dir.data(), nullptr, nullptr, nullptr) != 0) {
        std::cerr << "Error: " << altb_last_error() << std::endl;
        return;
    }
    delete lr.hAngularDist; delete lr.hDirectionZ;
    lr.hAngularDist = new TH1D("hAngularDist", "Angular Distribution of Exiting Rays;Angle from normal (degrees);Count", 180, -90, 90);
    lr.hDirectionZ = new TH1D("hDirectionZ", "Z Direction Component;Z;Count", 100, -1, 1);
    int fluxCount = 0;
    double exitPortZ = -100 * cm;
    for (int i = 0; i < n; ++i) {
        const double* lastPoint = &pos[(size_t)i * 3];
        const double* direction = &dir[(size_t)i * 3];
        if (lastPoint[2] < exitPortZ) {
            fluxCount++;
            double norm = std::sqrt(direction[0] * direction[0] + direction[1] * direction[1] + direction[2] * direction[2]);
            double dx = direction[0] / norm, dz = direction[2] / norm;
            lr.hDirectionZ->Fill(dz);
            double theta = std::copysign(std::acos(dz) * 180.0 / M_PI, dx);        // TMath::Sign(acos(dz)*180/pi, dx) (:94)
            if (std::isfinite(theta)) lr.hAngularDist->Fill(theta, 1.0);
        }
    }
    lr.fluxCount = fluxCount;
    std::cout << "Flux of rays through the exit port: " << fluxCount << std::endl;
}

// ------------------------------------------------------------------ C entry points (ctypes / tests / other hosts)
extern "C" {
void altbm_release() { ctx_cache().release(); }       // free the cached library contexts (device buffers, NCCL communicators)
int altbm_set(const char* key, double v) {
    Settings& S = settings();
    std::string k = key;
    if (k == "rays_per_position") S.rays_per_position = (int)v;
    else if (k == "traceonce_rays") S.traceonce_rays = (int)v;
    else if (k == "n_theta_bins") S.n_theta_bins = (int)v;
    else if (k == "n_phi_bins") S.n_phi_bins = (int)v;
    else if (k == "nrays_macro") S.nrays_macro = (int)v;
    else if (k == "sweep_rays") S.sweep_rays = (int)v;
    else if (k == "sweep_dtheta") S.sweep_dtheta = v;
    else if (k == "distribution_rays") S.distribution_rays = (int)v;
    else if (k == "nonlambertian_rays") S.nonlambertian_rays = (int)v;
    else if (k == "nonlambertian_posthoc") S.nonlambertian_posthoc = (int)v;
    else if (k == "contract") S.contract = (int)v;
    else if (k == "next_ray_id") S.next_ray_id = (uint64_t)v;
    else if (k == "advance_ray_ids") S.advance_ray_ids = (int)v;
    else if (k == "seed") S.seed = (uint64_t)v;
    else if (k == "traceonce_as_shipped") S.traceonce_as_shipped = (int)v;
    else if (k == "verbose") S.verbose = (int)v;
    else return -1;
    return 0;
}
void altbm_set_output_dir(const char* d) { settings().output_dir = d ? d : ""; }
const char* altbm_last_csv() { return last_run().csv_path.c_str(); }
long long altbm_last_count(const char* what) {
    std::string w = what; LastRun& r = last_run();
    if (w == "fluxCount") return r.fluxCount;
    if (w == "totalHitRays") return r.totalHitRays;
    if (w == "exitedRays") return r.exitedRays;
    if (w == "n_bounces") return r.n_bounces;
    if (w == "first_ray_id") return (long long)r.first_ray_id;
    if (w == "next_ray_id") return (long long)settings().next_ray_id;
    return -1;
}
double altbm_last_hist(const char* name, int bin) {
    LastRun& r = last_run(); std::string n = name;
    if (n == "hDirectionZ" && r.hDirectionZ) return r.hDirectionZ->GetBinContent(bin);
    if (n == "hAngularDist" && r.hAngularDist) return r.hAngularDist->GetBinContent(bin);
    return -1;
}
double altbm_last_fluxmap(int bx, int by) { return last_run().fluxMap ? last_run().fluxMap->GetBinContent(bx, by) : -1; }
int altbm_write_traceonce_csv(const char* path, const uint64_t* counts, int n, int nTheta, int nPhi, double thetaMax, double sx, double sy,
                              double sz, double dx, double dy, double dz, long long nExit, double rayTime, double sweepTime, double totalTime) {
    return writeTraceOnceFile(path, "2025-04-02 14:00:40", counts, n, nTheta, nPhi, thetaMax, sx, sy, sz, dx, dy, dz, nExit, rayTime, sweepTime,
                              totalTime) ? 0 : -1;
}
int altbm_write_perposition_csv(const char* path, int twofold, const uint64_t* counts, int n, int nTheta, int nPhi, double thetaMax, double sx,
                                double sy, double sz, double dx, double dy, double dz, double realTime) {
    return writePerPositionFile(path, "2025-04-01 01:42:14", twofold != 0, counts, n, nTheta, nPhi, thetaMax, sx, sy, sz, dx, dy, dz, realTime,
                                nullptr) ? 0 : -1;
}
const char* altbm_unique_filename(const char* base) { static std::string s; s = getUniqueFilename(base); return s.c_str(); }
void altbm_sweepDetector(int notify, const char* f, int t, double sx, double sy, double sz, double dx, double dy, double dz, double th) {
    fluxAtObserverOptimize::sweepDetector(notify != 0, f, t, sx, sy, sz, dx, dy, dz, th);
}
void altbm_sweepDetectorTwofold(int notify, const char* f, int t, double sx, double sy, double sz, double dx, double dy, double dz, double th) {
    fluxAtObserverFast::sweepDetectorTwofold(notify != 0, f, t, sx, sy, sz, dx, dy, dz, th);
}
void altbm_sweepDetectorTraceOnce(int notify, const char* f, int t, double sx, double sy, double sz, double dx, double dy, double dz, double th) {
    fluxAtObserverFast::sweepDetectorTraceOnce(notify != 0, f, t, sx, sy, sz, dx, dy, dz, th);
}
void altbm_sweepSeriesFast() { fluxAtObserverFast::sweepSeries(); }
void altbm_sweepSeriesOptimize() { fluxAtObserverOptimize::sweepSeries(); }
void altbm_fluxAtObserver_sweepDetector() { fluxAtObserver::sweepDetector(); }
void altbm_nonLambertianFlux_sweepDetector() { nonLambertianFlux::sweepDetector(); }
void altbm_makeIntegratingSphereNRays() { makeIntegratingSphereNRays(); }
void altbm_integratingSphereDetectorSweep() { integratingSphereDetectorSweep(); }
void altbm_distributionSphereDetectorSweep() { distributionSphereDetectorSweep(); }
}
