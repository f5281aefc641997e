// rootshim TH1/TH1D.  ROOT behaviour restated [from memory, SURVEY.md A.6]:
// TAxis::FindBin: x<xmin -> 0 (underflow); !(x<xmax) -> nbins+1 (overflow);
// else 1+int(nbins*(x-xmin)/(xmax-xmin)).  Fill(x,w) adds w to that bin.
// Integral() sums bins 1..nbins.
#ifndef ROOTSHIM_TH1D_h
#define ROOTSHIM_TH1D_h
#include <string>
#include <vector>
#include "TObject.h"
class TH1 : public TObject {
public:
    TH1(const char* name, const char* title, int nb, double lo, double hi)
        : fName(name), fTitle(title), fN(nb), fLo(lo), fHi(hi),
          fContent(nb + 2, 0.0) {}
    int FindBin(double x) const {
        if (x < fLo) return 0;
        if (!(x < fHi)) return fN + 1;
        return 1 + int(fN * (x - fLo) / (fHi - fLo));
    }
    int Fill(double x) { return Fill(x, 1.0); }
    int Fill(double x, double w) {
        int b = FindBin(x);
        fContent[b] += w;
        return b;
    }
    double GetBinContent(int b) const { return fContent[b]; }
    void SetBinContent(int b, double v) { fContent[b] = v; }
    int GetNbinsX() const { return fN; }
    void Reset() { fContent.assign(fN + 2, 0.0); }
    double Integral() const {
        double s = 0.0;
        for (int b = 1; b <= fN; ++b) s += fContent[b];
        return s;
    }
    virtual TObject* Clone(const char* name = "") const {
        TH1* h = new TH1(*this);
        h->fName = name;
        return h;
    }
    // TH1::Add(h1, c1) [from memory, ROOT 6 TH1.cxx]: every cell, under- and
    // overflow included, content += c1 * h1.content.  Sumw2() only enables the
    // error array, which nothing on this path reads.
    bool Add(const TH1* h, double c = 1.0) {
        for (size_t b = 0; b < fContent.size(); ++b) fContent[b] += c * h->fContent[b];
        return true;
    }
    void Sumw2(bool = true) {}
    void SetName(const char* n) { fName = n; }
    void SetLineColor(int) {}
    void SetTitle(const char* t) { fTitle = t; }
private:
    std::string fName, fTitle;
    int fN;
    double fLo, fHi;
    std::vector<double> fContent;
};
class TH1D : public TH1 {
public:
    TH1D(const char* name, const char* title, int nb, double lo, double hi)
        : TH1(name, title, nb, lo, hi) {}
    virtual TObject* Clone(const char* name = "") const {
        TH1D* h = new TH1D(*this);
        h->SetName(name);
        return h;
    }
};
#endif
