// root_dump_tape.C -- FOR A MACHINE THAT HAS ROOT + ROBAST (this repository's build and test boxes do not; the file is
// not compiled or run by anything here).  SURVEY.md 8f-4 / DESIGN.md "parity unpinned": the reference never dumped its
// initial rays or random draws, so bit-level parity against ROBAST cannot be pinned in this environment.  This macro is
// the missing half: it traces N rays of the fluxAtObserverFast.C scene with ROBAST and writes, per ray,
//   * every random number ROBAST drew, tagged by the TRandom call that produced it (U = Uniform/Rndm, G = Gaus), in order;
//   * every polyline point, the final status and direction.
// From that dump (a) the per-bounce draw ORDER of AOpticsManager::TraceNonSequential is read off (SURVEY appendix A.3
// assumes u_abs, u_psi, g, u_r, u_phi), (b) a tape in the format of include/altair_b200.h:altb_replay is built
// (8 f32 per surface hit: u_abs, u_r, u_phi, u_sel, u_psi, g0, g1, 0) and (c) altb_replay / altb_trace_paths are compared
// with ROBAST's own hit points ray by ray.
//
//   root -l -b -q 'root_dump_tape.C(1000, 170., "tape_dump.txt")'
#include "TGeoBBox.h"
#include "TGeoSphere.h"
#include "TRandom3.h"

#include "ABorderSurfaceCondition.h"
#include "AMirror.h"
#include "AOpticsManager.h"
#include "ARay.h"

#include <fstream>
#include <iomanip>
#include <vector>

class LoggingRandom : public TRandom3 {
public:
    struct Draw { char tag; double value; };
    std::vector<Draw> log;
    bool inside = false;      // Gaus/Uniform call Rndm internally: log only the outermost call
    LoggingRandom(UInt_t seed = 4357) : TRandom3(seed) {}
    Double_t Rndm() override {
        Double_t v = TRandom3::Rndm();
        if (!inside) log.push_back({'U', v});
        return v;
    }
    Double_t Uniform(Double_t x1 = 1) override { inside = true; Double_t v = TRandom3::Uniform(x1); inside = false; log.push_back({'U', v / x1}); return v; }
    Double_t Uniform(Double_t x1, Double_t x2) override {
        inside = true; Double_t v = TRandom3::Uniform(x1, x2); inside = false; log.push_back({'U', (v - x1) / (x2 - x1)}); return v;
    }
    Double_t Gaus(Double_t mean = 0, Double_t sigma = 1) override {
        inside = true; Double_t v = TRandom3::Gaus(mean, sigma); inside = false; log.push_back({'G', sigma != 0 ? (v - mean) / sigma : 0}); return v;
    }
};

void root_dump_tape(int n = 1000, double thetaMax = 170., const char* out = "tape_dump.txt") {
    const double cm = AOpticsManager::cm(), nm = AOpticsManager::nm();
    AOpticsManager* manager = new AOpticsManager("manager", "spherical shell");      // fluxAtObserverFast.C:192-230
    manager->SetLimit(50000);
    TGeoBBox* box = new TGeoBBox("box", 300 * cm, 300 * cm, 300 * cm);
    AOpticalComponent* world = new AOpticalComponent("world", box);
    manager->SetTopVolume(world);
    TGeoSphere* sphere = new TGeoSphere("sphereWithExitPort", 100.1 * cm, 101 * cm, 0., thetaMax);
    AMirror* mirror = new AMirror("mirror", sphere);
    mirror->SetReflectance(0.99);
    ABorderSurfaceCondition* condition = new ABorderSurfaceCondition(world, mirror);
    condition->EnableLambertian(true);
    condition->SetGaussianRoughness(0.01);
    world->AddNode(mirror, 1);
    manager->SetNsegments(100);
    manager->CloseGeometry();

    LoggingRandom* rng = new LoggingRandom(4357);
    delete gRandom;
    gRandom = rng;
    std::ofstream f(out);
    f << std::setprecision(17);
    f << "# ray <i> start x y z dx dy dz | draws: <tag value>... | points: x y z ... | status dir\n";
    for (int i = 0; i < n; ++i) {
        rng->log.clear();
        ARay* ray = new ARay(i, 660 * nm, -60 * cm, 0 * cm, -75 * cm, 0, 5, 0, 0);
        manager->TraceNonSequential(*ray);                                           // single-ray overload: one thread, ordered draws
        f << "ray " << i << " start -60 0 -75 5 0 0\n draws " << rng->log.size();
        for (auto& d : rng->log) f << " " << d.tag << " " << d.value;
        f << "\n points " << ray->GetNpoints();
        for (int k = 0; k < ray->GetNpoints(); k++) {
            Double_t p[4];
            ray->GetPoint(k, p);                                                     // ARay::GetPoint(Int_t, Double_t*) : x y z t
            f << " " << p[0] / cm << " " << p[1] / cm << " " << p[2] / cm;
        }
        Double_t d[3];
        ray->GetDirection(d);
        f << "\n status " << (ray->IsExited() ? 1 : ray->IsAbsorbed() ? 2 : ray->IsSuspended() ? 3 : 0) << " dir " << d[0] << " " << d[1] << " " << d[2] << "\n";
        delete ray;
    }
    f.close();
}
