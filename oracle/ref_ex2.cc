// oracle/ref_ex2.cc -- second translation unit of oracle/_ref/libsmcmc_ref.so:
// the reference's example2/FakeLikelihood.H (+ SystematicCorrection.H,
// Simulated.H, FakeData.H), included UNMODIFIED from /root/reference.  The
// example2 classes carry the same names and include guards as the ones of
// example/ (FakeLikelihood, SystematicCorrection, Simulated, FakeData), so they
// live in their own translation unit and inside `namespace ex2`; every header
// they include themselves is included first, at global scope.
// TEST INFRASTRUCTURE ONLY (see chain_api.h).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <TH1D.h>
#include <TRandom.h>
#include <TRandom3.h>
#include <TTree.h>
#include <TFile.h>
#include <TMatrixD.h>
#include <TVectorD.h>
#include <TMatrixDSymEigen.h>
#include <TDecompChol.h>

#include "chain_api.h"

#include "TSimpleMCMC.H"
using namespace sMCMC;  // example2/ predates the namespace

namespace ex2 {
#include "example2/FakeLikelihood.H"
}

extern "C" {

void* ref_ex2_create(void) {
    ex2::FakeLikelihood* like = new ex2::FakeLikelihood;
    like->DataClose = like->DataSeparated = like->DataDecayTag = 0;
    return like;
}

void ref_ex2_destroy(void* h) { delete static_cast<ex2::FakeLikelihood*>(h); }

// The tail of FakeLikelihood::Init (example2/FakeLikelihood.H:143-165) with the
// caller's events and data histograms instead of freshly generated ones.
int ref_ex2_set(void* h, const orc_event* ev, long n, const double* data150) {
    ex2::FakeLikelihood* like = static_cast<ex2::FakeLikelihood*>(h);
    static_assert(sizeof(ex2::Simulated::Event) == sizeof(orc_event), "event layout");
    like->SimulatedSample.resize(n);
    std::memcpy((void*)like->SimulatedSample.data(), ev, sizeof(orc_event) * n);
    TH1D* close = new TH1D("DataClose", "", 50, 0.0, 500.0);
    TH1D* separated = new TH1D("DataSeparated", "", 50, 0.0, 500.0);
    TH1D* tag = new TH1D("DataDecayTag", "", 50, 0.0, 500.0);
    for (int b = 0; b < 50; ++b) {
        close->SetBinContent(b + 1, data150[b]);
        separated->SetBinContent(b + 1, data150[50 + b]);
        tag->SetBinContent(b + 1, data150[100 + b]);
    }
    like->DataClose = close;
    like->DataSeparated = separated;
    like->DataDecayTag = tag;
    TH1D* blank = new TH1D("blank", "", 50, 0.0, 500.0);
    like->SimulatedSeparated = (TH1D*)blank->Clone("simSep");
    like->SimulatedSeparatedSignal = (TH1D*)blank->Clone("simSeparatedSig");
    like->SimulatedSeparatedBackground = (TH1D*)blank->Clone("simSepBkgd");
    like->SimulatedClose = (TH1D*)blank->Clone("simClose");
    like->SimulatedCloseSignal = (TH1D*)blank->Clone("simCloseSig");
    like->SimulatedCloseBackground = (TH1D*)blank->Clone("simCloseBkgd");
    like->SimulatedDecayTag = (TH1D*)blank->Clone("simDecayTag");
    like->SimulatedDecayTagSignal = (TH1D*)blank->Clone("simDecayTagSig");
    like->SimulatedDecayTagBackground = (TH1D*)blank->Clone("simDecayTagBkgd");
    return 0;
}

double ref_ex2_llh(void* h, const double* x, int n) {
    ex2::FakeLikelihood* like = static_cast<ex2::FakeLikelihood*>(h);
    Vector p(x, x + n);
    return (*like)(p);
}

// The renormalised expectation (Close, Separated, DecayTag) at x.
int ref_ex2_hist(void* h, const double* x, int n, double* out150) {
    ex2::FakeLikelihood* like = static_cast<ex2::FakeLikelihood*>(h);
    std::vector<double> p(x, x + n);
    like->ResetHistograms();
    like->FillHistograms(p);
    for (int b = 0; b < 50; ++b) {
        out150[b] = like->SimulatedClose->GetBinContent(b + 1);
        out150[50 + b] = like->SimulatedSeparated->GetBinContent(b + 1);
        out150[100 + b] = like->SimulatedDecayTag->GetBinContent(b + 1);
    }
    return 0;
}

// example2's own toy generators (Simulated::MakeSample, FakeData::FillData) run
// under a seeded shim generator; tests/test_synth.py compares their moments with
// smcmc_b200.synth.
long ref_ex2_generate(unsigned long seed, int dataSignal, int dataBackground, double oversample,
                      orc_event* out, long capacity, double* data150) {
    TRandom3 rng(seed);
    TRandom* old = gRandom;
    gRandom = &rng;
    std::streambuf* oldBuf = std::cout.rdbuf(0);
    ex2::FakeLikelihood like;
    like.Init(dataSignal, dataBackground, oversample);
    std::cout.rdbuf(oldBuf);
    gRandom = old;
    long n = (long)like.SimulatedSample.size();
    if (out && n <= capacity) std::memcpy(out, like.SimulatedSample.data(), sizeof(orc_event) * n);
    if (data150) {
        for (int b = 0; b < 50; ++b) {
            data150[b] = like.DataClose->GetBinContent(b + 1);
            data150[50 + b] = like.DataSeparated->GetBinContent(b + 1);
            data150[100 + b] = like.DataDecayTag->GetBinContent(b + 1);
        }
    }
    return n;
}

}  // extern "C"
