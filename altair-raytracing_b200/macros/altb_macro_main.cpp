// altb_macro -- command-line runner for the macro mirror, the stand-in for `root -l <macro>.C`:
//   altb_macro fluxAtObserverFast.C sweepDetectorTraceOnce [notify saveFolder threads srcX srcY srcZ dirX dirY dirZ thetaMax]
//   altb_macro fluxAtObserverFast.C sweepSeries
//   altb_macro makeIntegratingSphereNRays.C            (runs the function named like the file, as ROOT does)
// Settings: --set key=value (see altb_macros::Settings), --out DIR.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "altb_macros.h"

extern "C" int altbm_set(const char* key, double v);

static double num(const std::vector<std::string>& a, size_t i, double def) { return i < a.size() ? atof(a[i].c_str()) : def; }

int main(int argc, char** argv) {
    std::vector<std::string> a;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--set") && i + 1 < argc) {
            std::string kv = argv[++i];
            size_t eq = kv.find('=');
            if (eq == std::string::npos || altbm_set(kv.substr(0, eq).c_str(), atof(kv.c_str() + eq + 1)) != 0) {
                std::cerr << "unknown setting " << kv << std::endl;
                return 2;
            }
        } else if (!strcmp(argv[i], "--out") && i + 1 < argc) altb_macros::settings().output_dir = argv[++i];
        else a.push_back(argv[i]);
    }
    if (a.empty()) {
        std::cerr << "usage: altb_macro <macro.C> [function] [args...] [--set key=value] [--out DIR]" << std::endl;
        return 2;
    }
    std::string file = a[0];
    size_t slash = file.find_last_of('/');
    if (slash != std::string::npos) file = file.substr(slash + 1);
    if (file.size() > 2 && file.substr(file.size() - 2) == ".C") file = file.substr(0, file.size() - 2);
    std::string fn = a.size() > 1 ? a[1] : file;
    std::vector<std::string> r(a.begin() + (a.size() > 1 ? 2 : 1), a.end());
    bool notify = r.size() > 0 ? atoi(r[0].c_str()) != 0 : true;
    std::string folder = r.size() > 1 ? r[1] : "results";
    int threads = (int)num(r, 2, -1);
    double sx = num(r, 3, -60), sy = num(r, 4, 0), sz = num(r, 5, -80), dx = num(r, 6, 5), dy = num(r, 7, 2), dz = num(r, 8, 0), th = num(r, 9, 170.);
    if (file == "fluxAtObserverOptimize") {
        if (fn == "sweepDetector") fluxAtObserverOptimize::sweepDetector(notify, folder.c_str(), threads, sx, sy, sz, dx, dy, dz, th);
        else if (fn == "sweepSeries" || fn == file) fluxAtObserverOptimize::sweepSeries();
        else goto bad;
    } else if (file == "fluxAtObserverFast") {
        if (fn == "sweepDetector") fluxAtObserverFast::sweepDetector(notify, folder.c_str(), threads, sx, sy, sz, dx, dy, dz, th);
        else if (fn == "sweepDetectorTwofold") fluxAtObserverFast::sweepDetectorTwofold(notify, folder.c_str(), threads, sx, sy, sz, dx, dy, dz, th);
        else if (fn == "sweepDetectorTraceOnce") fluxAtObserverFast::sweepDetectorTraceOnce(notify, folder.c_str(), threads, sx, sy, sz, dx, dy, dz, th);
        else if (fn == "sweepSeries" || fn == file) fluxAtObserverFast::sweepSeries();
        else goto bad;
    } else if (file == "fluxAtObserver") fluxAtObserver::sweepDetector();
    else if (file == "nonLambertianFlux") nonLambertianFlux::sweepDetector();
    else if (file == "makeIntegratingSphereNRays") makeIntegratingSphereNRays();
    else if (file == "integratingSphereDetectorSweep") integratingSphereDetectorSweep();
    else if (file == "distributionSphereDetectorSweep") distributionSphereDetectorSweep();
    else goto bad;
    return 0;
bad:
    std::cerr << "no such macro/function: " << file << " " << fn << std::endl;
    return 2;
}
