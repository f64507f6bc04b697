// altb_th.h -- the few TH1D/TH2D calls the reference macros make (TH2D fluxMap(180,0,90; 90,0,360),
// SetBinContent(i+1,j+1,f), Fill), as a dense array, for hosts without ROOT.  With ROOT present, compile the
// macros with -DALTB_WITH_ROOT and these names resolve to ROOT's classes instead.
#pragma once
#ifndef ALTB_WITH_ROOT
#include <string>
#include <vector>

class TH1D {
public:
    TH1D(const char* name, const char* title, int nx, double xlo, double xhi)
        : fName(name), fTitle(title), fN(nx), fLo(xlo), fHi(xhi), fC(nx + 2, 0.0) {}
    int Fill(double x, double w = 1.0) { int b = FindBin(x); fC[b] += w; fEntries += 1; return b; }
    int FindBin(double x) const { if (x < fLo) return 0; if (!(x < fHi)) return fN + 1; return 1 + (int)((x - fLo) / (fHi - fLo) * fN); }
    double GetBinContent(int b) const { return fC[b]; }
    void SetBinContent(int b, double v) { fC[b] = v; }
    double GetBinCenter(int b) const { return fLo + (b - 0.5) * (fHi - fLo) / fN; }
    int GetNbinsX() const { return fN; }
    double GetEntries() const { return fEntries; }
    double Integral() const { double s = 0; for (int b = 1; b <= fN; b++) s += fC[b]; return s; }
    const char* GetName() const { return fName.c_str(); }
private:
    std::string fName, fTitle; int fN; double fLo, fHi; std::vector<double> fC; double fEntries = 0;
};

class TH2D {
public:
    TH2D(const char* name, const char* title, int nx, double xlo, double xhi, int ny, double ylo, double yhi)
        : fName(name), fTitle(title), fNx(nx), fNy(ny), fXlo(xlo), fXhi(xhi), fYlo(ylo), fYhi(yhi),
          fC((size_t)(nx + 2) * (ny + 2), 0.0) {}
    void SetBinContent(int bx, int by, double v) { fC[(size_t)by * (fNx + 2) + bx] = v; }
    double GetBinContent(int bx, int by) const { return fC[(size_t)by * (fNx + 2) + bx]; }
    int Fill(double x, double y, double w = 1.0) {
        int bx = x < fXlo ? 0 : (!(x < fXhi) ? fNx + 1 : 1 + (int)((x - fXlo) / (fXhi - fXlo) * fNx));
        int by = y < fYlo ? 0 : (!(y < fYhi) ? fNy + 1 : 1 + (int)((y - fYlo) / (fYhi - fYlo) * fNy));
        fC[(size_t)by * (fNx + 2) + bx] += w;
        return by * (fNx + 2) + bx;
    }
    int GetNbinsX() const { return fNx; }
    int GetNbinsY() const { return fNy; }
    double Integral() const { double s = 0; for (int y = 1; y <= fNy; y++) for (int x = 1; x <= fNx; x++) s += GetBinContent(x, y); return s; }
    const char* GetName() const { return fName.c_str(); }
private:
    std::string fName, fTitle; int fNx, fNy; double fXlo, fXhi, fYlo, fYhi; std::vector<double> fC;
};
#else
#include "TH1D.h"
#include "TH2D.h"
#endif
