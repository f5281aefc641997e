// oracle/smcmc_oracle.cc -- the CPU restatement ("port") of the hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file restates, in plain C++ over flat
// arrays and with NO dependency on ROOT or on the reference tree, the
// algorithm of the path BASELINE.json names: the Metropolis step
// (TSimpleMCMC.H:370-496), the adaptive proposal (TSimpleMCMC.H:640-1831) and
// the likelihood functors of SURVEY.md section 8a.  Each function cites the
// reference lines it follows.  It is pinned by tests/test_oracle_vs_ref.py,
// which runs it side by side with oracle/_ref/libsmcmc_ref.so (the
// reference's own headers compiled unmodified) on identical injected draws
// and requires bit-identical chains, and by the golden vectors in
// tests/golden/ that were produced by that reference build.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the resulting library.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "chain_api.h"
#include "smcmc_rng.h"

namespace {

std::string gLastError;

// ---------------------------------------------------------------------------
// Third-party (ROOT) primitives, restated [SURVEY.md A.6].
// ---------------------------------------------------------------------------

// TDecompChol::Decompose: a = U^T U, column ordered, upper triangle only.
// Returns false on a non-positive pivot.  `a` and `u` are n x n row-major.
bool CholeskyUpper(const std::vector<double>& a, std::vector<double>& u, int n) {
    u = a;
    for (int c = 0; c < n; ++c) {
        double pivot = u[(size_t)c * n + c];
        for (int r = 0; r < c; ++r) pivot -= u[(size_t)r * n + c] * u[(size_t)r * n + c];
        if (pivot <= 0.0) return false;
        pivot = std::sqrt(pivot);
        u[(size_t)c * n + c] = pivot;
        for (int j = c + 1; j < n; ++j) {
            double v = u[(size_t)c * n + j];
            for (int r = 0; r < c; ++r) v -= u[(size_t)r * n + j] * u[(size_t)r * n + c];
            u[(size_t)c * n + j] = v / pivot;
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) u[(size_t)i * n + j] = 0.0;
    return true;
}

// TMatrixDSymEigen: eigenvalues descending, eigenvectors in columns.  Cyclic
// Jacobi, the same procedure as oracle/rootshim/TMatrixD.h (the true ROOT
// routine is tridiagonalisation + QL; eigenvectors are only defined up to
// sign/order, so this stage is property-tested).
void SymEigen(const std::vector<double>& in, int n, std::vector<double>& vec,
              std::vector<double>& val) {
    std::vector<double> a(in);
    std::vector<double> v((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) v[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) off += a[(size_t)p * n + q] * a[(size_t)p * n + q];
        if (!(off > 1e-300)) break;
        for (int p = 0; p < n; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double apq = a[(size_t)p * n + q];
                if (apq == 0.0) continue;
                double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
                    a[(size_t)k * n + p] = c * akp - s * akq;
                    a[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
                    a[(size_t)p * n + k] = c * apk - s * aqk;
                    a[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    double vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q];
                    v[(size_t)k * n + p] = c * vkp - s * vkq;
                    v[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
        }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        return a[(size_t)x * n + x] > a[(size_t)y * n + y];
    });
    val.assign(n, 0.0);
    vec.assign((size_t)n * n, 0.0);
    for (int c = 0; c < n; ++c) {
        val[c] = a[(size_t)order[c] * n + order[c]];
        for (int r = 0; r < n; ++r) vec[(size_t)r * n + c] = v[(size_t)r * n + order[c]];
    }
}

// TAxis::FindBin for the example's 50 bins on [0,500).
inline int FindBin50(double x) {
    if (x < 0.0) return 0;
    if (!(x < 500.0)) return 51;
    return 1 + int(50 * (x - 0.0) / (500.0 - 0.0));
}

// ---------------------------------------------------------------------------
// Likelihoods.
// ---------------------------------------------------------------------------
struct Likelihood {
    int kind = ORC_LLH_UNIT_GAUSS;
    int dim = 0;
    std::vector<double> error;          // ORC_LLH_DUMMY: n x n row-major
    std::vector<orc_event> events;      // ORC_LLH_FAKE
    double data[150];                   // close, separated, decay-tag
    double exposure = 1.0;
    double sim[3][52];                  // filled by FillFake

    // example/SystematicCorrection.H:50-79
    static double InvariantMass(const orc_event& e, const double* p) {
        double mass = e.Mass;
        if (e.Type < 0) return mass;
        double nomLogMass = std::log(e.TrueMass);
        double nomLogSigma = std::log(e.TrueMass + e.TrueMassSigma);
        nomLogSigma = nomLogSigma - nomLogMass;
        double logMass = std::log(mass);
        double logSigma = (logMass - nomLogMass) / nomLogSigma;
        double scale = p[2] / 10.0;
        double width = std::exp(p[3] / 10.0);
        double skew = 0.3 * std::erf(p[4] / 10.0);
        skew = std::exp(logSigma * skew);
        logMass = nomLogMass + (logMass - nomLogMass) * skew;
        logMass = nomLogMass + (logMass - nomLogMass) * width;
        logMass = logMass + scale;
        return std::exp(logMass);
    }
    // example/SystematicCorrection.H:35-48
    static double Separation(const orc_event& e, const double* p) {
        if (e.Type < 0) return e.Separation;
        double scale = 0.0;
        if (e.Type == 0) scale += p[5];
        if (e.Type > 0) scale += p[6];
        scale = std::exp(scale / 10.0);
        return e.Separation * scale;
    }
    // example/SystematicCorrection.H:81-117
    double EventWeight(const orc_event& e, const double* p) const {
        double weight = 1.0;
        if (e.Type < 0) return weight;
        if (e.Type == 0) weight *= std::exp(p[0] / 10.0);
        else weight *= std::exp(p[1] / 10.0);
        const double trueFakes = 0.05;
        double fakes = std::tan(M_PI * (trueFakes - 0.5));
        fakes += p[7];
        fakes = std::atan(fakes) / M_PI + 0.5;
        if (e.Type == 0) {
            if (e.MuDk > 0) weight *= fakes / trueFakes;
            else weight *= (1.0 - fakes) / (1.0 - trueFakes);
        }
        const double trueEff = 0.5;
        double eff = std::tan(M_PI * (trueEff - 0.5));
        eff += p[8];
        eff = std::atan(eff) / M_PI + 0.5;
        if (e.Type > 0) {
            if (e.MuDk > 0) weight *= eff / trueEff;
            else weight *= (1.0 - eff) / (1.0 - trueEff);
        }
        weight *= exposure;
        return weight;
    }
    // example/FakeLikelihood.H:188-216
    void FillFake(const double* p) {
        std::memset(sim, 0, sizeof(sim));
        for (size_t i = 0; i < events.size(); ++i) {
            const orc_event& e = events[i];
            double mass = InvariantMass(e, p);
            double sep = Separation(e, p);
            double w = EventWeight(e, p);
            if (mass > 500.0) continue;
            if (mass < 0.0) continue;
            if (sep < 0.0) continue;
            int h = 1;                       // separated
            if (e.MuDk > 0) h = 2;           // decay tag
            else if (sep < 100.0) h = 0;     // close
            sim[h][FindBin50(mass)] += w;
        }
    }
    // The same loop, but counting events per weight class instead of summing
    // weights: slot layout of the device count table (signal untagged 0-99,
    // signal tagged 100-149, background untagged 150-249, background tagged
    // 250-299, data-typed 300-449; inside a block: Close bins, Separated bins).
    void CountFake(const double* p, uint32_t* out450) {
        std::memset(out450, 0, 450 * sizeof(uint32_t));
        for (size_t i = 0; i < events.size(); ++i) {
            const orc_event& e = events[i];
            double mass = InvariantMass(e, p);
            double sep = Separation(e, p);
            if (mass > 500.0) continue;
            if (mass < 0.0) continue;
            if (sep < 0.0) continue;
            int bin = FindBin50(mass);
            if (bin < 1 || bin > 50) continue;
            int h = 1;
            if (e.MuDk > 0) h = 2;
            else if (sep < 100.0) h = 0;
            int slot;
            if (e.Type < 0) slot = 300 + h * 50;
            else {
                int base = (e.Type == 0) ? 0 : 150;
                slot = (h == 2) ? base + 100 : base + h * 50;
            }
            out450[slot + bin - 1] += 1;
        }
    }
    // example/FakeLikelihood.H:47-81
    double EvalFake(const double* p) {
        FillFake(p);
        double logLikelihood = 0.0;
        for (int h = 0; h < 3; ++h) {
            for (int b = 1; b <= 50; ++b) {
                double d = data[h * 50 + b - 1];
                double mc = sim[h][b];
                if (mc < 0.001) mc = 0.001;
                double v = d - mc;
                if (d > 0.0) v += d * std::log(mc / d);
                logLikelihood += v;
            }
        }
        return logLikelihood;
    }

    // ---- example2 ---------------------------------------------------------
    // example2/SystematicCorrection.H:75-117: no event-count factor, no exposure
    static double EventWeight2(const orc_event& e, const double* p) {
        double weight = 1.0;
        if (e.Type < 0) return weight;
        const double trueFakes = 0.05;
        double fakes = std::tan(M_PI * (trueFakes - 0.5));
        fakes += p[7];
        fakes = std::atan(fakes) / M_PI + 0.5;
        if (e.Type == 0) {
            if (e.MuDk > 0) weight *= fakes / trueFakes;
            else weight *= (1.0 - fakes) / (1.0 - trueFakes);
        }
        const double trueEff = 0.5;
        double eff = std::tan(M_PI * (trueEff - 0.5));
        eff += p[8];
        eff = std::atan(eff) / M_PI + 0.5;
        if (e.Type > 0) {
            if (e.MuDk > 0) weight *= eff / trueEff;
            else weight *= (1.0 - eff) / (1.0 - trueEff);
        }
        return weight;
    }
    // example2/FakeLikelihood.H:222-289: six histograms (signal | background x
    // Close, Separated, DecayTag), their integrals, and the renormalised
    // expectation left in sim[h][1..50].  Mass, separation and the cuts are
    // those of example/ (identical functions, example2/SystematicCorrection.H:30-73).
    void FillFake2(const double* p) {
        double part[2][3][52];
        std::memset(part, 0, sizeof(part));
        for (size_t i = 0; i < events.size(); ++i) {
            const orc_event& e = events[i];
            double mass = InvariantMass(e, p);
            double sep = Separation(e, p);
            double w = EventWeight2(e, p);
            if (mass > 500.0) continue;
            if (mass < 0.0) continue;
            if (sep < 0.0) continue;
            int h = 1;
            if (e.MuDk > 0) h = 2;
            else if (sep < 100.0) h = 0;
            part[e.Type == 0 ? 0 : 1][h][FindBin50(mass)] += w;      // IsSignal: Type == 0
        }
        double norm[2];
        for (int k = 0; k < 2; ++k) {
            double integral[3];
            for (int h = 0; h < 3; ++h) {                           // TH1::Integral: bins 1..50
                double t = 0.0;
                for (int b = 1; b <= 50; ++b) t += part[k][h][b];
                integral[h] = t;
            }
            double t = integral[2];                                 // DecayTag, Close, Separated :266-275
            t += integral[0];
            t += integral[1];
            norm[k] = p[k] / t;
        }
        std::memset(sim, 0, sizeof(sim));
        for (int h = 0; h < 3; ++h)
            for (int b = 0; b < 52; ++b) {                          // TH1::Add, every cell :280-287
                sim[h][b] += norm[0] * part[0][h][b];
                sim[h][b] += norm[1] * part[1][h][b];
            }
    }
    // example2/FakeLikelihood.H:58-118
    double EvalFake2(const double* p) {
        FillFake2(p);
        double logLikelihood = 0.0;
        for (int h = 0; h < 3; ++h) {
            for (int b = 1; b <= 50; ++b) {
                double d = data[h * 50 + b - 1];
                double mc = sim[h][b];
                if (mc < 0.001) mc = 0.001;
                double v = d - mc;
                if (d > 0.0) v += d * std::log(mc / d);
                logLikelihood += v;
            }
        }
        double v = p[0];
        if (v < 0.0) logLikelihood -= 10.0 + std::abs(logLikelihood);
        v = p[1];
        if (v < 0.0) logLikelihood -= 10.0 + std::abs(logLikelihood);
        v = p[6] / 5.0;
        logLikelihood -= 0.5 * v * v;
        v = p[7] / 1.0;
        logLikelihood -= 0.5 * v * v;
        v = p[8] / 1.0;
        logLikelihood -= 0.5 * v * v;
        return logLikelihood;
    }

    double operator()(const double* x) {
        switch (kind) {
        case ORC_LLH_UNIT_GAUSS: {           // TSimpleMCMC.H:113-119
            double s = 0.0;
            for (int i = 0; i < dim; ++i) s += -0.5 * x[i] * x[i];
            return s;
        }
        case ORC_LLH_DUMMY: {                // TDummyLogLikelihood.H:21-31
            double s = 0.0;
            for (int i = 0; i < dim; ++i)
                for (int j = 0; j < dim; ++j)
                    s -= 0.5 * x[i] * error[(size_t)j * dim + i] * x[j];
            return s;
        }
        case ORC_LLH_HORRIFIC: {             // THorrificLogLikelihood.H:26-38
            const double sigma = 0.01;
            double s = 0.0;
            for (int i = 0; i < dim; ++i) {
                if (std::abs(x[i]) > 1.0) return -1E+30;
                s += x[i];
            }
            double naturalSigma = std::sqrt(dim * 4.0 / 12.0);
            s /= naturalSigma;
            s = -0.5 * s * s / sigma / sigma;
            return s;
        }
        case ORC_LLH_ASYM: {                 // TAsymLogLikelihood.H:20-31
            double s = 0.0;
            for (int i = 0; i < dim; ++i) {
                double a = x[i];
                if (a < 0.0) a *= 100.0;
                else a *= -1.0;
                s += a;
            }
            return s;
        }
        case ORC_LLH_FAKE:
            return EvalFake(x);
        case ORC_LLH_FAKE2:
            return EvalFake2(x);
        case ORC_LLH_UNBINNED:
            return EvalUnbinned(x);
        case ORC_LLH_CONSTRAINED: {          // example4/TConstrainedLikelihood.H:26-46, priors of Init() :55-110
            double logLikelihood = 0.0;
            double sum = 0.0;
            for (int i = 0; i < dim; ++i) sum += x[i];
            sum = (sum - 1902.0) / 16.0;
            logLikelihood -= 0.5 * sum * sum;
            for (int i = 0; i < dim; ++i) {
                const double expected = (i < 24) ? 76.0 : 80.0;
                const double prior = (i < 24) ? 76.0 * 0.08 : 2.0;
                double v = x[i] - expected;
                v /= prior;
                logLikelihood -= 0.5 * v * v;
            }
            return logLikelihood;
        }
        case ORC_LLH_HARD: {                 // THardLogLikelihood.H:57-69
            double s = 0.0;
            for (int i = 0; i < dim - 1; ++i) {
                double a = (1.0 - x[i]);
                double b = x[i + 1] - x[i] * x[i];
                s -= a * a + 100.0 * b * b;
            }
            return s;
        }
        }
        return std::numeric_limits<double>::quiet_NaN();
    }

    // The unbinned mixture likelihood of BASELINE.json configs[4].  It has no
    // counterpart in the reference; the definition is in
    // include/smcmc_b200.h (SMCMC_LLH_UNBINNED) and reuses the reference's
    // per-event corrections: the corrected log-mass of
    // SystematicCorrection::InvariantMass (example/SystematicCorrection.H:50-79)
    // and the two hypothesis weights of EventWeight (:81-117).
    double EvalUnbinned(const double* p) const {
        const double scale = p[2] / 10.0;
        const double width = std::exp(p[3] / 10.0);
        const double skewc = 0.3 * std::erf(p[4] / 10.0);
        double fakes = std::tan(M_PI * (0.05 - 0.5));
        fakes += p[7];
        fakes = std::atan(fakes) / M_PI + 0.5;
        double eff = std::tan(M_PI * (0.5 - 0.5));
        eff += p[8];
        eff = std::atan(eff) / M_PI + 0.5;
        const double wSig = 1.0 * std::exp(p[0] / 10.0), wBkg = 1.0 * std::exp(p[1] / 10.0);
        // log weight of the signal / background hypothesis, untagged and tagged
        const double lws[2] = {std::log(wSig * ((1.0 - fakes) / (1.0 - 0.05))), std::log(wSig * (fakes / 0.05))};
        const double lwb[2] = {std::log(wBkg * ((1.0 - eff) / (1.0 - 0.5))), std::log(wBkg * (eff / 0.5))};
        const double mu = std::log(135.0), sig = std::log(1.3), tau = 500.0;
        const double ca = -std::log(sig * std::sqrt(2.0 * M_PI)), cb = -std::log(tau);
        double sum = 0.0;
        for (size_t i = 0; i < events.size(); ++i) {
            const orc_event& e = events[i];
            const double nomLog = std::log(e.TrueMass);
            const double nomLogSigma = std::log(e.TrueMass + e.TrueMassSigma) - nomLog;
            const double d = std::log(e.Mass) - nomLog;
            const double logSigma = d / nomLogSigma;
            const double skew = std::exp(logSigma * skewc);
            double lm = nomLog + d * skew;
            lm = nomLog + (lm - nomLog) * width;
            lm = lm + scale;
            const int tag = e.MuDk > 0 ? 1 : 0;
            const double z = (lm - mu) / sig;
            const double a = lws[tag] + ca - 0.5 * z * z - lm;         // log(w_s phi_s(m)), lognormal around 135
            const double b = lwb[tag] + cb - std::exp(lm) / tau;       // log(w_b phi_b(m)), exponential, tau = 500
            const double hi = a > b ? a : b, lo = a > b ? b : a;
            sum += hi + std::log1p(std::exp(lo - hi));
        }
        return sum;
    }
};

// ---------------------------------------------------------------------------
// TProposeAdaptiveStep (TSimpleMCMC.H:640-1977), flat-array restatement.
// ---------------------------------------------------------------------------
struct Proposal {
    int n = 0;
    uint64_t seed = 0;
    uint32_t chain = 0;
    std::vector<double> lastPoint, center, centerChange, cov, decomp;
    std::vector<int> type;
    std::vector<double> param1, param2;
    struct Corr { int d1, d2; double c; };
    std::vector<Corr> corr;
    double lastValue = 0.0;
    double centerTrials = 0.0, covTrials = 0.0, covDeweight = 0.5;
    bool covFrozen = false;
    double covWindow = -1;
    int trials = 0, successes = 0, nextUpdate = -1;
    double acceptance = 0.0, acceptanceTrials = 0.0, acceptanceDeweight = 0.5;
    double acceptanceWindow = -1, rigidity = 2.0, target = -1;
    double sigma = 0.0, sigmaTrace = 0.0;
    double maxCorrelation = 1.0 - std::sqrt(std::numeric_limits<double>::epsilon());
    bool initialized = false;
    bool failed = false;

    void SetDim(int d) {                      // :786-795
        if (!lastPoint.empty()) return;
        n = d;
        lastPoint.assign(d, 0.0);
        type.assign(d, 0);
        param1.assign(d, 0.0);
        param2.assign(d, 0.0);
    }
    double Trace() const {                    // :961-967
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += cov[(size_t)i * n + i];
        return t;
    }
    void SetCorrelation(int d1, int d2, double c) {   // :883-904
        if (d1 == d2) return;
        if (c < -maxCorrelation) c = -maxCorrelation;
        if (c > maxCorrelation) c = maxCorrelation;
        corr.push_back(Corr{d1, d2, c});
    }

    // :1009-1390.  Returns false when the reference would throw.
    bool UpdateProposal(bool fromReset) {
        double trace = Trace();
        if (trace <= 0) { gLastError = "Invalid trace"; return false; }      // :1024-1028
        sigma = sigma * std::sqrt(sigmaTrace / trace);                       // :1042
        sigmaTrace = trace;
        double maxUp = (double)((size_t)n * (size_t)n);                      // :1050-1052
        double up = 0.5 * successes;
        nextUpdate = (int)(acceptanceWindow + maxUp - maxUp / (up + 1.0));
        if (covDeweight > 0.0) {                                             // :1056-1067
            if (covDeweight > 1.0) covDeweight = 1.0;
            double w = 1.0 - covDeweight;
            covTrials = std::max(1.0, w * covTrials);
            covTrials = std::min(covTrials, w * covWindow);
            centerTrials = std::max(1.0, w * centerTrials);
            centerTrials = std::min(centerTrials, w * covWindow);
        }
        if (acceptanceDeweight > 0.0) {                                      // :1081-1086
            if (acceptanceDeweight > 1.0) acceptanceDeweight = 1.0;
            double w = 1.0 - acceptanceDeweight;
            acceptanceTrials = std::max(1.0, w * acceptanceTrials);
            acceptanceTrials = std::min(acceptanceTrials, w * acceptanceWindow);
        }
        const double minVar = std::numeric_limits<double>::epsilon();
        std::vector<double> u;
        if (CholeskyUpper(cov, u, n)) { decomp = u; return true; }           // :1103-1120
        // Condition the variances, :1134-1183.
        for (int i = 0; i < n; ++i) {
            double expected = 1.0;
            if (type[i] == 0) {
                if (param1[i] > 0) expected = param1[i];
            } else {
                expected = param2[i];
                expected -= param1[i];
                expected = expected * expected / 12.0;
            }
            double& v = cov[(size_t)i * n + i];
            if (!std::isfinite(v)) v = expected;
            if (v < 0.0) v = minVar * expected;
            if (v < minVar * expected) v = minVar * expected;
            if (v < minVar) v = minVar;
        }
        // Condition the correlations, :1187-1217.
        for (int i = 0; i < n; ++i) {
            for (int j = i + 1; j < n; ++j) {
                double c = cov[(size_t)i * n + j];
                c /= std::sqrt(cov[(size_t)i * n + i]);
                c /= std::sqrt(cov[(size_t)j * n + j]);
                if (!std::isfinite(c)) c = 0.0;
                if (std::abs(c) > maxCorrelation) c = (c > 0.0) ? maxCorrelation : -maxCorrelation;
                double v = c;
                v *= std::sqrt(cov[(size_t)i * n + i]);
                v *= std::sqrt(cov[(size_t)j * n + j]);
                cov[(size_t)i * n + j] = v;
                cov[(size_t)j * n + i] = v;
            }
        }
        if (CholeskyUpper(cov, u, n)) { decomp = u; return true; }           // :1220-1239
        // Eigen-decomposition fallback, :1252-1321.
        {
            std::vector<double> sym((size_t)n * n), vec, val;
            for (int i = 0; i < n; ++i)
                for (int j = i; j < n; ++j)
                    sym[(size_t)j * n + i] = sym[(size_t)i * n + j] = cov[(size_t)i * n + j];
            SymEigen(sym, n, vec, val);
            double eigenSum = 0.0;
            for (int i = 0; i < n; ++i) {
                if (val[i] < 0.0) continue;
                eigenSum += val[i];
            }
            double minAxis = 1.0 - maxCorrelation;
            if (minAxis < minVar) minAxis = minVar;
            minAxis = minAxis * val[0];
            for (int i = 0; i < n; ++i) {
                double rms = std::sqrt(std::max(minAxis, val[i]));
                for (int j = 0; j < n; ++j) decomp[(size_t)i * n + j] = rms * vec[(size_t)j * n + i];
            }
            if (eigenSum > 1E-6) return true;
        }
        // Emergency shrink, :1335-1377.
        double step = std::numeric_limits<double>::epsilon();
        for (int i = 0; i < n; ++i) step = std::max(step, cov[(size_t)i * n + i]);
        step *= 1E-4;
        double dec = 1.0;
        for (int trial = 0; trial < 10; ++trial) {
            dec *= 0.84;
            for (int i = 0; i < n; ++i) {
                cov[(size_t)i * n + i] += step;
                for (int j = i + 1; j < n; ++j) {
                    double v = dec * cov[(size_t)i * n + j];
                    cov[(size_t)j * n + i] = v;
                    cov[(size_t)i * n + j] = v;
                }
            }
            if (CholeskyUpper(cov, u, n)) { decomp = u; return true; }
        }
        if (fromReset) { gLastError = "Decomposition of user correlations failed"; return false; }
        return ResetProposal();                                             // :1389
    }

    // :1396-1494
    bool ResetProposal() {
        trials = 0;
        successes = 0;
        if (sigma < 0.01 * std::sqrt(1.0 / n)) sigma = std::sqrt(1.0 / n);
        decomp.resize((size_t)n * n, 0.0);
        cov.resize((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i) {
            for (int j = i; j < n; ++j) {
                if (i == j && type[i] == 0 && param1[i] > 0) cov[(size_t)i * n + i] = param1[i];
                else if (i == j && type[i] == 1) {
                    double delta = param1[i];
                    delta -= param2[i];
                    cov[(size_t)i * n + i] = delta * delta / 12.0;
                } else if (i == j) cov[(size_t)i * n + i] = 1.0;
                else cov[(size_t)i * n + j] = cov[(size_t)j * n + i] = 0.0;
            }
        }
        for (size_t k = 0; k < corr.size(); ++k) {                           // :1445-1457
            const Corr& c = corr[k];
            if (c.d1 == c.d2) continue;
            double v1 = cov[(size_t)c.d1 * n + c.d1];
            double v2 = cov[(size_t)c.d2 * n + c.d2];
            double v = c.c * std::sqrt(v1) * std::sqrt(v2);
            cov[(size_t)c.d1 * n + c.d2] = v;
            cov[(size_t)c.d2 * n + c.d1] = v;
        }
        sigmaTrace = Trace();                                                // :1460
        int minWindow = 100 + 4 * n;                                         // :1468-1476
        if (covWindow < minWindow) {
            covWindow = n;
            covWindow *= n;
            covWindow *= n;
            covWindow += minWindow;
            double r = std::numeric_limits<double>::epsilon();
            covWindow = std::min(covWindow, std::sqrt(1.0 / r));
        }
        if (target < 0.0) { gLastError = "Target acceptance not initialized"; return false; }
        acceptance = target;                                                 // :1481-1482
        acceptanceTrials = std::min(10.0, 0.5 * acceptanceWindow);
        center = lastPoint;                                                  // :1484-1491
        centerChange.assign(n, 0.0);
        centerTrials = std::max(centerTrials, 1.0);
        return UpdateProposal(true);
    }

    // :1679-1714
    bool InitializeState(const double* current, double value) {
        if (initialized) return true;
        initialized = true;
        if (lastPoint.empty()) SetDim(n);
        lastValue = value;
        std::copy(current, current + n, lastPoint.begin());
        if (acceptanceWindow < 0) acceptanceWindow = std::pow(1.0 * n, 1.5) + 1000;
        nextUpdate = (int)acceptanceWindow;
        if (target < 1E-4) {
            if (n > 4) target = 0.234;
            else target = 0.44;
        }
        return ResetProposal();
    }

    // :1721-1831
    bool UpdateState(const double* current, double value) {
        if (!InitializeState(current, value)) return false;
        ++trials;
        bool accepted = false;
        if (value != lastValue || current[0] != lastPoint[0]) accepted = true;
        if (accepted) ++successes;
        acceptance *= acceptanceTrials;
        if (accepted) acceptance = acceptance + 1.0;
        acceptance /= acceptanceTrials + 1.0;
        acceptanceTrials = std::min(acceptanceWindow, acceptanceTrials + 1.0);
        if (rigidity < 500.0 && rigidity > 0.0) {                            // :1745-1762
            double accSigma = target * (1.0 - target);
            accSigma = std::sqrt(accSigma / acceptanceWindow);
            if (std::abs(acceptance - target) < accSigma) {
                rigidity += 0.5 * rigidity / acceptanceWindow;
                rigidity = std::min(200.0, rigidity);
            }
            if (std::abs(acceptance - target) > 4.0 * accSigma) {
                rigidity -= 1.618 * 0.5 * rigidity / acceptanceWindow;
                rigidity = std::max(2.0, rigidity);
            }
        }
        if (rigidity > 0 && rigidity < 100.0) {                              // :1771-1776
            sigma *= std::pow(acceptance / target,
                              std::min(1.0 / 500.0, 1.0 / (rigidity * acceptanceWindow)));
        }
        for (int i = 0; i < n; ++i) {                                        // :1780-1788
            centerChange[i] = center[i];
            center[i] *= centerTrials;
            center[i] += current[i];
            center[i] /= centerTrials + 1;
            centerChange[i] = center[i] - centerChange[i];
        }
        centerTrials = std::min(covWindow, centerTrials + 1.0);
        if (!covFrozen) {                                                    // :1795-1820
            for (int i = 0; i < n; ++i) {
                for (int j = 0; j < i + 1; ++j) {
                    double v = cov[(size_t)i * n + j];
                    double r = (current[i] - center[i]) * (current[j] - center[j]);
                    v *= covTrials;
                    v += r;
                    v /= covTrials + 1.0;
                    cov[(size_t)i * n + j] = v;
                    cov[(size_t)j * n + i] = v;
                }
            }
            covTrials = std::min(covWindow, covTrials + 1.0);
        }
        if (accepted && (--nextUpdate) < 1) {                                // :1824-1826
            if (!UpdateProposal(false)) return false;
        }
        lastValue = value;
        std::copy(current, current + n, lastPoint.begin());
        return true;
    }

    std::vector<double> forcedStep;          // fForcedStep  :811-818
    int scanDimension = -1;                  // fScanDimension :820-830
    uint32_t drawsUsed = 0;                  // gRandom calls of the last operator()

    // operator(), :659-725
    bool Propose(double* proposal, const double* current, double value, uint32_t step) {
        if (!forcedStep.empty()) {                                           // :671-678
            std::copy(forcedStep.begin(), forcedStep.end(), proposal);
            forcedStep.clear();
            drawsUsed = 0;
            return true;
        }
        const int scan = scanDimension;
        if (scan >= 0 && scan < n) {                                         // :685-704
            std::copy(current, current + n, proposal);
            drawsUsed = 1;
            if (type[scan] == 1) {
                double u = smcmc_uniform(seed, chain, step, 0u, SMCMC_STREAM_STEP);
                proposal[scan] = param1[scan] + (param2[scan] - param1[scan]) * u;
                return true;
            }
            double sig = 1.0;
            if (param1[scan] > 0) sig = std::sqrt(param1[scan]);
            double g = smcmc_normal(seed, chain, step, 0u, SMCMC_STREAM_STEP);
            proposal[scan] = center[scan] + sig * g;                         // TRandom::Gaus(mean, sigma)
            return true;
        }
        drawsUsed = (uint32_t)n;
        if (!UpdateState(current, value)) return false;
        std::copy(current, current + n, proposal);
        for (int i = 0; i < n; ++i) {
            if (type[i] == 1) {
                double u = smcmc_uniform(seed, chain, step, (uint32_t)i, SMCMC_STREAM_STEP);
                proposal[i] = param1[i] + (param2[i] - param1[i]) * u;       // TRandom::Uniform(a,b)
                continue;
            }
            double g = smcmc_normal(seed, chain, step, (uint32_t)i, SMCMC_STREAM_STEP);
            double r = 0.0 + 1.0 * g;                                        // TRandom::Gaus(0,1)
            for (int j = 0; j < n; ++j) {
                if (type[j] == 1) continue;
                proposal[j] += sigma * r * decomp[(size_t)i * n + j];
            }
        }
        return true;
    }
};

// ---------------------------------------------------------------------------
// TProposeVAATStep (TProposeVAATStep.H:22-307): adaptive variable-at-a-time
// proposal.  Draws of a step, in the order the reference calls gRandom: when
// the index queue is empty n uniforms for the shuffle (:186-189), then one draw
// for the proposed coordinate (:60-78); TSimpleMCMC's accept uniform follows.
// ---------------------------------------------------------------------------
struct VaatProposal {
    int n = 0;
    uint64_t seed = 0;
    uint32_t chain = 0;
    std::vector<int> type;
    std::vector<double> param1, param2;
    std::vector<int> nextIndex;
    std::vector<double> acceptance, sigma;
    std::vector<int> acceptanceTrials;
    double lastValue = 0.0;
    int trials = 0, successes = 0;
    int acceptanceWindow = -1;                 // an int in the reference (:281)
    int lastIndex = -1;
    double rigidity = 2.0, target = 0.44;
    bool initialized = false;
    uint32_t slot = 0;                         // draws consumed by the current step

    void SetDim(int d) {                       // :82-96
        n = d;
        type.assign(d, 0);
        param1.assign(d, 0.0);
        param2.assign(d, 0.0);
        acceptance.assign(d, 0.0);
        acceptanceTrials.assign(d, 0);
        sigma.assign(d, 2.34);
    }
    void InitializeState(double value) {       // :193-209
        if (initialized) return;
        initialized = true;
        lastValue = value;
        acceptanceWindow = 100;
    }
    void UpdateState(double value) {           // :216-254
        InitializeState(value);
        ++trials;
        bool accepted = false;
        if (value != lastValue) accepted = true;
        if (accepted) ++successes;
        lastValue = value;
        if (lastIndex < 0) return;
        ++acceptanceTrials[lastIndex];
        acceptance[lastIndex] *= 1.0 * std::min(acceptanceWindow, acceptanceTrials[lastIndex]);
        if (accepted) acceptance[lastIndex] += 1.0;
        acceptance[lastIndex] /= 1.0 + 1.0 * std::min(acceptanceWindow, acceptanceTrials[lastIndex]);
        if (acceptanceTrials[lastIndex] > 0.1 * acceptanceWindow && rigidity > 0 && rigidity < 100.0) {
            double v = sigma[lastIndex];
            v *= std::pow(acceptance[lastIndex] / target,
                          std::min(1.0 / 500.0, 1.0 / (rigidity * acceptanceWindow)));
            sigma[lastIndex] = std::max(v, 1.0E-4);
        }
    }
    void UpdateProposal(uint32_t step) {       // :176-190
        if (!nextIndex.empty()) return;
        nextIndex.resize(n);
        lastIndex = -1;
        for (int i = 0; i < n; ++i) nextIndex[i] = i;
        for (int i = 0; i < n; ++i) {
            double u = 1.0 * smcmc_uniform(seed, chain, step, slot++, SMCMC_STREAM_STEP);   // TRandom::Uniform()
            std::size_t sw = (std::size_t)(nextIndex.size() * u);
            if (sw >= nextIndex.size()) sw = nextIndex.size() - 1;     // u < 1, but n*u can round up to n
            std::swap(nextIndex[i], nextIndex[sw]);
        }
    }
    void Propose(double* proposal, const double* current, double value, uint32_t step) {   // :40-80
        slot = 0;
        UpdateState(value);
        std::copy(current, current + n, proposal);
        UpdateProposal(step);
        lastIndex = nextIndex.back();
        nextIndex.pop_back();
        if (type[lastIndex] == 1) {
            double u = smcmc_uniform(seed, chain, step, slot++, SMCMC_STREAM_STEP);
            proposal[lastIndex] = param1[lastIndex] + (param2[lastIndex] - param1[lastIndex]) * u;
            return;
        }
        double expectedVariance = 1.0;
        if (type[lastIndex] == 0 && param1[lastIndex] > 0) expectedVariance = param1[lastIndex];
        double g = smcmc_normal(seed, chain, step, slot++, SMCMC_STREAM_STEP);
        double r = 0.0 + expectedVariance * g;                                   // TRandom::Gaus(0, expectedVariance)
        proposal[lastIndex] = current[lastIndex] + sigma[lastIndex] * r;
    }
    double MeanSigma() const {                 // GetSigma :166-173
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += sigma[i];
        return n ? t / n : 0.0;
    }
};

// ---------------------------------------------------------------------------
// TSimpleMCMC (TSimpleMCMC.H:185-590)
// ---------------------------------------------------------------------------
struct OrcChain {
    Likelihood like;
    Proposal prop;
    VaatProposal vprop;
    bool vaat = false;      // TSimpleMCMC<L, TProposeVAATStep>
    int n = 0;
    uint32_t step = 0;
    std::vector<double> accepted, proposed, trial;
    double acceptedLlh = 0.0, proposedLlh = 0.0;
    double stepRMS = 0.0;
    int stepRMSTrials = 0, stepRMSWindow = 1000;
    int totalSteps = 0, llhCalls = 0;

    double Eval(const double* x) { ++llhCalls; return like(x); }       // :538-541

    int Start(const double* x0) {                                      // :246-276
        proposed.assign(x0, x0 + n);
        accepted.assign(x0, x0 + n);
        trial.assign(x0, x0 + n);
        proposedLlh = Eval(proposed.data());
        if (!std::isfinite(proposedLlh) || proposedLlh < -0.999999E+10) return 0;
        acceptedLlh = proposedLlh;
        if (vaat) {
            vprop.InitializeState(acceptedLlh);
            return 1;
        }
        if (!prop.InitializeState(accepted.data(), acceptedLlh)) return -1;
        return 1;
    }

    // :370-496.  Returns 1 accepted, 0 rejected, -1 error.
    int Step(int metropolis) {
        if (proposed.empty() || accepted.empty()) { gLastError = "Uninitialized starting point"; return -1; }
        ++totalSteps;
        uint32_t acceptSlot = (uint32_t)n;
        if (vaat) {
            vprop.Propose(proposed.data(), accepted.data(), acceptedLlh, step++);
            acceptSlot = vprop.slot;
        } else {
            if (!prop.Propose(proposed.data(), accepted.data(), acceptedLlh, step++)) return -1;
            acceptSlot = prop.drawsUsed;                               // the accept draw is the next gRandom call
        }
        if (stepRMSWindow > 0) {                                       // :391-406
            double sqr = 0.0;
            for (int i = 0; i < n; ++i) {
                trial[i] = proposed[i] - accepted[i];
                sqr += trial[i] * trial[i];
            }
            double ms = stepRMS * stepRMS;
            ms *= stepRMSTrials;
            ms += sqr;
            ms /= stepRMSTrials + 1.0;
            stepRMSTrials = std::min(stepRMSWindow, stepRMSTrials + 1);
            stepRMS = std::sqrt(ms);
        }
        proposedLlh = Eval(proposed.data());                           // :410
        if (metropolis == 2) {                                         // :414-426
            accepted = proposed;
            acceptedLlh = proposedLlh;
            return 1;
        }
        if (!std::isfinite(proposedLlh) || proposedLlh < -0.999999E+30) return 0;   // :432-436
        double delta = proposedLlh - acceptedLlh;                      // :441-463
        if (delta < 0.0) {
            if (metropolis == 1) return 0;
            double u = 1.0 * smcmc_uniform(prop.seed, prop.chain, step - 1, acceptSlot, SMCMC_STREAM_STEP);
            double t = std::log(u);
            if (delta < t) return 0;
        }
        acceptedLlh = proposedLlh;                                     // :484-491
        accepted = proposed;
        return 1;
    }
};

// What SaveStep(true) leaves in the last tree entry (TSimpleMCMC.H:208-216,
// :1631-1652), and Restore()/RestoreState() (:282-352, :1501-1610) from it.
struct SavedState {
    std::vector<double> accepted, center, covPacked;
    double llh, stepRMS, acceptance, acceptanceTrials, sigma, centerTrials, covTrials;
    int totalSteps, trials, successes, nextUpdate;
};

SavedState SaveFull(const OrcChain& c) {
    SavedState s;
    const Proposal& p = c.prop;
    s.accepted = c.accepted;
    s.llh = c.acceptedLlh;
    s.totalSteps = c.totalSteps;
    s.stepRMS = c.stepRMS;
    s.trials = p.trials;
    s.successes = p.successes;
    s.nextUpdate = p.nextUpdate;
    s.acceptance = p.acceptance;
    s.acceptanceTrials = p.acceptanceTrials;
    s.sigma = p.sigma;
    s.center = p.center;
    s.centerTrials = p.centerTrials;
    s.covTrials = p.covTrials;
    for (int i = 0; i < c.n; ++i)
        for (int j = 0; j < i + 1; ++j) s.covPacked.push_back(p.cov[(size_t)i * c.n + j]);
    return s;
}

bool RestoreFrom(OrcChain& c, const SavedState& s) {
    const int n = c.n;
    c.totalSteps = s.totalSteps;                                     // :320-331
    c.acceptedLlh = s.llh;
    c.stepRMS = s.stepRMS;
    c.accepted = s.accepted;
    c.proposed = s.accepted;
    c.trial = s.accepted;
    c.proposedLlh = c.Eval(c.proposed.data());                       // :335-345
    if (std::abs(c.proposedLlh - c.acceptedLlh) > 1E-4) c.acceptedLlh = c.proposedLlh;
    Proposal& p = c.prop;                                            // RestoreState :1501-1610
    p.initialized = true;
    p.lastValue = c.acceptedLlh;
    p.lastPoint = c.accepted;
    p.trials = s.trials;
    p.successes = s.successes;
    p.nextUpdate = s.nextUpdate;
    p.acceptance = s.acceptance;
    p.acceptanceTrials = s.acceptanceTrials;
    p.sigma = s.sigma;
    p.center = s.center;
    p.centerTrials = s.centerTrials;
    if (s.covPacked.size() != (size_t)n * (n + 1) / 2) { gLastError = "Past the end of the covariance"; return false; }
    size_t k = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i + 1; ++j) {
            p.cov[(size_t)i * n + j] = s.covPacked[k];
            p.cov[(size_t)j * n + i] = s.covPacked[k];
            ++k;
        }
    p.sigmaTrace = p.Trace();
    p.covTrials = s.covTrials;
    return p.UpdateProposal(false);                                  // :1607
}

OrcChain* H(void* h) { return static_cast<OrcChain*>(h); }

}  // namespace

extern "C" {

const char* orc_last_error(void) { return gLastError.c_str(); }

void* orc_chain_create(int kind, int dim, uint64_t seed, uint32_t chain) {
    if (kind < ORC_LLH_UNIT_GAUSS || kind > ORC_LLH_CONSTRAINED || dim < 1 || (kind == ORC_LLH_HARD && dim < 2) ||
        (kind == ORC_LLH_CONSTRAINED && dim != 25)) {
        gLastError = "bad likelihood kind or dimension";
        return 0;
    }
    if ((kind == ORC_LLH_FAKE || kind == ORC_LLH_FAKE2 || kind == ORC_LLH_UNBINNED) && dim != 9) { gLastError = "FakeLikelihood is 9-dim"; return 0; }
    OrcChain* c = new OrcChain;
    c->n = dim;
    c->like.kind = kind;
    c->like.dim = dim;
    c->prop.seed = seed;
    c->prop.chain = chain;
    c->prop.n = dim;
    c->prop.SetDim(dim);
    return c;
}

void* orc_chain_create_vaat(int kind, int dim, uint64_t seed, uint32_t chain) {
    OrcChain* c = static_cast<OrcChain*>(orc_chain_create(kind, dim, seed, chain));
    if (!c) return 0;
    c->vaat = true;
    c->vprop.seed = seed;
    c->vprop.chain = chain;
    c->vprop.SetDim(dim);
    return c;
}

int orc_chain_get_vaat(void* h, double* sigma, double* acceptance, int32_t* acceptanceTrials, int32_t* misc) {
    OrcChain* c = H(h);
    if (!c->vaat) { gLastError = "not a TProposeVAATStep chain"; return -1; }
    VaatProposal& p = c->vprop;
    if (sigma) std::copy(p.sigma.begin(), p.sigma.end(), sigma);
    if (acceptance) std::copy(p.acceptance.begin(), p.acceptance.end(), acceptance);
    if (acceptanceTrials) std::copy(p.acceptanceTrials.begin(), p.acceptanceTrials.end(), acceptanceTrials);
    if (misc) {
        misc[0] = p.trials;
        misc[1] = p.successes;
        misc[2] = p.lastIndex;
        misc[3] = (int32_t)p.nextIndex.size();
    }
    return 0;
}

void orc_chain_destroy(void* h) { delete H(h); }

int orc_chain_set_fake(void* h, const orc_event* ev, long n, const double* data150, double exposure) {
    Likelihood& l = H(h)->like;
    if (l.kind != ORC_LLH_FAKE && l.kind != ORC_LLH_FAKE2 && l.kind != ORC_LLH_UNBINNED) { gLastError = "not a FakeLikelihood chain"; return -1; }
    l.events.assign(ev, ev + n);
    if (data150) std::copy(data150, data150 + 150, l.data);
    l.exposure = exposure;
    return 0;
}

int orc_chain_set_error_matrix(void* h, const double* e, int n) {
    Likelihood& l = H(h)->like;
    if (l.kind != ORC_LLH_DUMMY || n != l.dim) { gLastError = "error matrix shape"; return -1; }
    l.error.assign(e, e + (size_t)n * n);
    return 0;
}

int orc_chain_set(void* h, int field, double v) {
    if (H(h)->vaat && field != ORC_SET_STEP_RMS_WINDOW) {
        VaatProposal& q = H(h)->vprop;
        if (field == ORC_SET_ACCEPTANCE_WINDOW) q.acceptanceWindow = (int)v;        // :136
        else if (field == ORC_SET_ACCEPTANCE_RIGIDITY) q.rigidity = v;              // :146
        else { gLastError = "TProposeVAATStep has no such setting"; return -1; }
        return 0;
    }
    Proposal& p = H(h)->prop;
    switch (field) {
    case ORC_SET_SIGMA: p.sigma = v; break;
    case ORC_SET_TARGET_ACCEPTANCE: p.target = v; break;
    case ORC_SET_ACCEPTANCE_WINDOW: p.acceptanceWindow = v; break;
    case ORC_SET_ACCEPTANCE_RIGIDITY: p.rigidity = v; break;
    case ORC_SET_ACCEPTANCE_DEWEIGHT: p.acceptanceDeweight = v; break;
    case ORC_SET_COVARIANCE_WINDOW: p.covWindow = (int)v; break;      // SetCovarianceWindow(int) :914
    case ORC_SET_COVARIANCE_DEWEIGHT: p.covDeweight = v; break;
    case ORC_SET_COVARIANCE_FROZEN: p.covFrozen = (v != 0.0); break;
    case ORC_SET_COVARIANCE_TRIALS: p.covTrials = v; break;
    case ORC_SET_CENTER_TRIALS: p.centerTrials = v; break;
    case ORC_SET_NEXT_UPDATE: p.nextUpdate = (int)v; break;           // int member :1930
    case ORC_SET_MAX_CORRELATION: p.maxCorrelation = v; break;
    case ORC_SET_STEP_RMS_WINDOW: H(h)->stepRMSWindow = (int)v; break;
    default: gLastError = "unknown field"; return -1;
    }
    return 0;
}

int orc_chain_set_gaussian(void* h, int d, double sigma) {            // :855-867
    Proposal& p = H(h)->prop;
    if (d < 0 || d >= p.n) return -1;
    if (H(h)->vaat) {                                                 // TProposeVAATStep.H:121-133: sigma itself
        H(h)->vprop.type[d] = 0;
        H(h)->vprop.param1[d] = sigma;
        return 0;
    }
    p.type[d] = 0;
    p.param1[d] = sigma * sigma;
    return 0;
}

int orc_chain_set_uniform(void* h, int d, double lo, double hi) {     // :833-848
    Proposal& p = H(h)->prop;
    if (d < 0 || d >= p.n) return -1;
    if (H(h)->vaat) {                                                 // TProposeVAATStep.H:99-115
        H(h)->vprop.type[d] = 1;
        H(h)->vprop.param1[d] = lo;
        H(h)->vprop.param2[d] = hi;
        return 0;
    }
    p.type[d] = 1;
    p.param1[d] = lo;
    p.param2[d] = hi;
    return 0;
}

int orc_chain_set_correlation(void* h, int d1, int d2, double c) {
    H(h)->prop.SetCorrelation(d1, d2, c);
    return 0;
}

int orc_chain_start(void* h, const double* x0) { return H(h)->Start(x0); }

int orc_chain_step(void* h, int nsteps, int metropolis, int32_t* accepted,
                   double* llhAccepted, double* llhProposed, double* x, double* sigma) {
    OrcChain* c = H(h);
    for (int s = 0; s < nsteps; ++s) {
        int r = c->Step(metropolis);
        if (r < 0) return -1;
        if (accepted) accepted[s] = r;
        if (llhAccepted) llhAccepted[s] = c->acceptedLlh;
        if (llhProposed) llhProposed[s] = c->proposedLlh;
        if (sigma) sigma[s] = c->vaat ? c->vprop.MeanSigma() : c->prop.sigma;
        if (x) std::copy(c->accepted.begin(), c->accepted.end(), x + (size_t)s * c->n);
    }
    return 0;
}

int orc_chain_update_proposal(void* h) { return H(h)->prop.UpdateProposal(false) ? 0 : -1; }
int orc_chain_reset_proposal(void* h) { return H(h)->prop.ResetProposal() ? 0 : -1; }

int orc_chain_get_state(void* h, double* s, double* accepted, double* center,
                        double* cov, double* decomp) {
    OrcChain* c = H(h);
    if (c->vaat) {
        VaatProposal& q = c->vprop;
        if (s) {
            for (int k = 0; k < ORC_ST_COUNT; ++k) s[k] = 0.0;
            double acc = 0.0;                                   // GetAcceptance :155-163
            for (int i = 0; i < q.n; ++i) acc += q.acceptance[i];
            s[ORC_ST_SIGMA] = q.MeanSigma();
            s[ORC_ST_ACCEPTANCE] = q.n ? acc / q.n : 0.0;
            s[ORC_ST_ACCEPTANCE_WINDOW] = q.acceptanceWindow;
            s[ORC_ST_ACCEPTANCE_RIGIDITY] = q.rigidity;
            s[ORC_ST_TRIALS] = q.trials;
            s[ORC_ST_SUCCESSES] = q.successes;
            s[ORC_ST_STEP_RMS] = c->stepRMS;
            s[ORC_ST_ACCEPTED_LLH] = c->acceptedLlh;
            s[ORC_ST_PROPOSED_LLH] = c->proposedLlh;
            s[ORC_ST_TOTAL_STEPS] = c->totalSteps;
            s[ORC_ST_LLH_CALLS] = c->llhCalls;
        }
        if (accepted) std::copy(c->accepted.begin(), c->accepted.end(), accepted);
        return 0;
    }
    Proposal& p = c->prop;
    const size_t nn = (size_t)c->n * c->n;
    if (s) {
        s[ORC_ST_SIGMA] = p.sigma;
        s[ORC_ST_ACCEPTANCE] = p.acceptance;
        s[ORC_ST_ACCEPTANCE_TRIALS] = p.acceptanceTrials;
        s[ORC_ST_ACCEPTANCE_WINDOW] = p.acceptanceWindow;
        s[ORC_ST_ACCEPTANCE_RIGIDITY] = p.rigidity;
        s[ORC_ST_TARGET_ACCEPTANCE] = p.target;
        s[ORC_ST_TRIALS] = p.trials;
        s[ORC_ST_SUCCESSES] = p.successes;
        s[ORC_ST_NEXT_UPDATE] = p.nextUpdate;
        s[ORC_ST_COVARIANCE_TRIALS] = p.covTrials;
        s[ORC_ST_COVARIANCE_WINDOW] = p.covWindow;
        s[ORC_ST_CENTER_TRIALS] = p.centerTrials;
        s[ORC_ST_COVARIANCE_TRACE] = p.cov.size() == nn ? p.Trace() : 0.0;
        s[ORC_ST_SIGMA_TRACE] = p.sigmaTrace;
        s[ORC_ST_STEP_RMS] = c->stepRMS;
        s[ORC_ST_ACCEPTED_LLH] = c->acceptedLlh;
        s[ORC_ST_PROPOSED_LLH] = c->proposedLlh;
        s[ORC_ST_TOTAL_STEPS] = c->totalSteps;
        s[ORC_ST_LLH_CALLS] = c->llhCalls;
    }
    if (accepted) std::copy(c->accepted.begin(), c->accepted.end(), accepted);
    if (center && p.center.size() == (size_t)c->n) std::copy(p.center.begin(), p.center.end(), center);
    if (cov && p.cov.size() == nn) std::copy(p.cov.begin(), p.cov.end(), cov);
    if (decomp && p.decomp.size() == nn) std::copy(p.decomp.begin(), p.decomp.end(), decomp);
    return 0;
}

double orc_chain_llh(void* h, const double* x) { return H(h)->like(x); }

int orc_chain_fake_hist(void* h, const double* x, double* out150) {
    Likelihood& l = H(h)->like;
    if (l.kind != ORC_LLH_FAKE && l.kind != ORC_LLH_FAKE2) { gLastError = "not a FakeLikelihood chain"; return -1; }
    if (l.kind == ORC_LLH_FAKE2) l.FillFake2(x);
    else l.FillFake(x);
    for (int hh = 0; hh < 3; ++hh)
        for (int b = 0; b < 50; ++b) out150[hh * 50 + b] = l.sim[hh][b + 1];
    return 0;
}

int orc_chain_force_step(void* h, const double* x) {
    OrcChain* c = H(h);
    c->prop.forcedStep.assign(x, x + c->n);
    return 0;
}
int orc_chain_set_scan(void* h, int dim) {
    OrcChain* c = H(h);
    if (c->prop.lastPoint.empty()) return 0;
    c->prop.scanDimension = (dim < 0 || dim >= c->n) ? -1 : dim;
    return 0;
}
int orc_chain_set_center(void* h, const double* v) {
    OrcChain* c = H(h);
    if ((int)c->prop.center.size() != c->n) return -1;
    std::copy(v, v + c->n, c->prop.center.begin());
    return 0;
}

int orc_chain_restore(void* h, void* source) {
    OrcChain* c = H(h);
    OrcChain* src = H(source);
    if (!RestoreFrom(*c, SaveFull(*src))) return -1;
    c->step = src->step;
    return 0;
}

// For symmetry with the reference-backed checker: saving is a no-op here, the
// full state is read from the source chain at restore time.
int orc_chain_save_step(void*) { return 0; }
int orc_chain_step_saved(void* h, int nsteps, int32_t* accepted) {
    return orc_chain_step(h, nsteps, 0, accepted, 0, 0, 0, 0);
}

int orc_chain_fake_counts(void* h, const double* x, uint32_t* out450) {
    Likelihood& l = H(h)->like;
    if (l.kind != ORC_LLH_FAKE && l.kind != ORC_LLH_FAKE2) { gLastError = "not a FakeLikelihood chain"; return -1; }
    l.CountFake(x, out450);
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// TSimpleHMC (TSimpleHMC.H:119-973), flat-array restatement.
// ---------------------------------------------------------------------------
namespace {

// TMatrixD::Invert as restated by the shim (Gauss-Jordan, partial pivoting).
void InvertInPlace(std::vector<double>& m, int n) {
    std::vector<double> a(m), inv((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = std::abs(a[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double v = std::abs(a[(size_t)i * n + k]);
            if (v > best) { best = v; piv = i; }
        }
        if (piv != k) {
            for (int j = 0; j < n; ++j) {
                std::swap(a[(size_t)k * n + j], a[(size_t)piv * n + j]);
                std::swap(inv[(size_t)k * n + j], inv[(size_t)piv * n + j]);
            }
        }
        double p = a[(size_t)k * n + k];
        for (int j = 0; j < n; ++j) {
            a[(size_t)k * n + j] /= p;
            inv[(size_t)k * n + j] /= p;
        }
        for (int i = 0; i < n; ++i) {
            if (i == k) continue;
            double f = a[(size_t)i * n + k];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) {
                a[(size_t)i * n + j] -= f * a[(size_t)k * n + j];
                inv[(size_t)i * n + j] -= f * inv[(size_t)k * n + j];
            }
        }
    }
    m.swap(inv);
}

struct OrcHmc {
    Likelihood like;
    bool withGradient = false;
    int n = 0;
    uint64_t seed = 0;
    uint32_t chain = 0, step = 0, slot = 0;
    // TSimpleHMC members
    int stepCount = 0, potentialCount = 0, gradientCount = 0, leapFrogSteps = 10;
    double alpha = 0.0, covWindow = 1000000;
    double currentAcceptance = 0, targetAcceptance = 0, meanEpsilon = 0, reversalLen = 0;
    std::vector<double> accepted, acceptedMomentum, proposed, proposedMomentum, central, average;
    double acceptedPotential = 0, proposedPotential = 0, centralPotential = 0;
    double averageTrials = 0, covTrials = 0, orbitLength = 0, estCovTrace = 0, curCovTrace = 0;
    std::vector<double> estCov, exxt, estErr;
    int stepsRemaining = 0, stepsSinceUpdate = 0;

    double Rndm() { return smcmc_uniform(seed, chain, step, slot++, SMCMC_STREAM_STEP); }
    double Gaus() { return 0.0 + 1.0 * smcmc_normal(seed, chain, step, slot++, SMCMC_STREAM_STEP); }

    double Potential(const std::vector<double>& x) {                  // :411-414
        ++potentialCount;
        return -like(x.data());
    }
    // user gradient of log(likelihood): TDummyLogLikelihood.H:34-42
    bool UserGradient(std::vector<double>& g, const std::vector<double>& p) {
        if (withGradient && like.kind == ORC_LLH_HARD) {              // THardLogLikelihood.H:72-91
            const double B = 100.0;
            g[0] = -2.0 * (1.0 - p[0]) - 4.0 * B * p[0] * (p[1] - p[0] * p[0]);
            for (int i = 1; i < n - 1; ++i) {
                g[i] = 2.0 * B * (p[i] - p[i - 1] * p[i - 1]);
                g[i] += -2.0 * (1.0 - p[i]);
                g[i] += -4.0 * B * p[i] * (p[i + 1] - p[i] * p[i]);
            }
            g[n - 1] = +2.0 * B * (p[n - 1] - p[n - 2] * p[n - 2]);
            for (int i = 0; i < n; ++i) g[i] = -g[i];
            return true;
        }
        if (!withGradient || like.kind != ORC_LLH_DUMMY) return false;
        for (int i = 0; i < n; ++i) {
            g[i] = 0.0;
            for (int j = 0; j < n; ++j) g[i] -= like.error[(size_t)i * n + j] * p[j];
        }
        return true;
    }
    void FiniteDifference(std::vector<double>& grad, const std::vector<double>& point) {   // :417-444
        std::vector<double> work(n);
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) work[j] = point[j];
            double du = 0.01;
            work[i] -= du;
            double u1 = Potential(work);
            work[i] += 2.0 * du;
            double u2 = Potential(work);
            grad[i] = 0.5 * (u2 - u1) / du;
        }
    }
    void Covariant(std::vector<double>& grad, const std::vector<double>& point) {          // :447-454
        for (int i = 0; i < n; ++i) {
            grad[i] = 0.0;
            for (int j = 0; j < n; ++j) grad[i] += estErr[(size_t)i * n + j] * (point[j] - average[j]);
        }
    }
    bool PotentialGradient(std::vector<double>& grad, const std::vector<double>& point, int type) {   // :467-532
        ++gradientCount;
        switch (type) {
        default:
        case 0:
        case 1:
            if (UserGradient(grad, point)) {
                for (int i = 0; i < n; ++i) grad[i] = -grad[i];
                return true;
            }
            FiniteDifference(grad, point);
            return true;
        case 2: Covariant(grad, point); return true;
        case 3: FiniteDifference(grad, point); return true;
        case 4:
            if (!UserGradient(grad, point)) { gLastError = "user gradient required"; return false; }
            for (int i = 0; i < n; ++i) grad[i] = -grad[i];
            return true;
        case 5:
            for (int i = 0; i < n; ++i) grad[i] = 0.0;
            return true;
        }
    }
    double Kinetic(const std::vector<double>& m) {                    // :535-542
        double ke = 0.0;
        for (int i = 0; i < n; ++i) { double p = m[i]; ke += p * p / 2.0; }
        return ke;
    }
    void ProposeMomentum(std::vector<double>& pNew, const std::vector<double>& m) {   // :554-570
        if (alpha >= 1.0) {
            alpha = std::max(1.0, alpha);
            for (int i = 0; i < n; ++i) pNew[i] = m[i] / alpha;
            return;
        }
        if (alpha < 0.0) alpha = 0.0;
        for (int i = 0; i < n; ++i) pNew[i] = alpha * m[i] + std::sqrt(1.0 - alpha * alpha) * Gaus();
    }
    int LeapFrog(std::vector<double>& qNew, std::vector<double>& pNew, const std::vector<double>& position,
                 double epsilon, int steps, int type) {               // :582-651
        qNew = position;
        std::vector<double> momentum = pNew, grad(n);
        int status = 1;
        if (steps < 1) {
            for (int j = 0; j < n; ++j) qNew[j] = qNew[j] + epsilon * (momentum[j] + pNew[j]) / 2.0;
            return status;
        }
        if (!PotentialGradient(grad, qNew, type)) return -1;
        for (int j = 0; j < n; ++j) pNew[j] = pNew[j] - epsilon * grad[j] / 2.0;
        for (int i = 0; i < steps - 1; ++i) {
            for (int j = 0; j < n; ++j) qNew[j] = qNew[j] + epsilon * pNew[j];
            if (!PotentialGradient(grad, qNew, type)) return -1;
            for (int j = 0; j < n; ++j) pNew[j] = pNew[j] - epsilon * grad[j];
            double inner = 0.0;
            for (int j = 0; j < n; ++j) inner += pNew[j] * momentum[j];
            if (inner >= 0.0) continue;
            status = 2;
        }
        for (int j = 0; j < n; ++j) qNew[j] = qNew[j] + epsilon * pNew[j];
        if (!PotentialGradient(grad, qNew, type)) return -1;
        for (int j = 0; j < n; ++j) pNew[j] = pNew[j] - epsilon * grad[j] / 2.0;
        return status;
    }
    void UpdateCovariance() {                                         // :665-695
        ++stepsSinceUpdate;
        --stepsRemaining;
        for (int i = 0; i < n; ++i) {
            double v = average[i];
            v *= averageTrials;
            v += accepted[i];
            v /= averageTrials + 1.0;
            average[i] = v;
        }
        averageTrials = std::min(covWindow, averageTrials + 1.0);
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < i + 1; ++j) {
                double v = exxt[(size_t)i * n + j];
                v *= covTrials;
                v += accepted[i] * accepted[j];
                v /= covTrials + 1.0;
                exxt[(size_t)i * n + j] = exxt[(size_t)j * n + i] = v;
                double c = v - average[i] * average[j];
                estCov[(size_t)i * n + j] = estCov[(size_t)j * n + i] = c;
            }
        }
        covTrials = std::min(covWindow, covTrials + 1.0);
    }
    void UpdateErrorMatrix() {                                        // :703-858
        if (!leapFrogSteps) return;
        if (covTrials < 2 * n) return;
        curCovTrace = 0.0;
        for (int i = 0; i < n; ++i) curCovTrace += std::abs(estCov[(size_t)i * n + i]);
        double change = std::abs(curCovTrace - estCovTrace);
        bool doIt = false;
        if (stepsRemaining < 0) doIt = true;
        if (stepsSinceUpdate > 2.0 * n && change > 0.01 * estCovTrace) doIt = true;
        if (!doIt) return;
        double aPot = Potential(average);                             // :729-738
        if (aPot < centralPotential) { central = average; centralPotential = aPot; }
        stepsRemaining = 2 * n + stepCount;                           // :758-759
        stepsSinceUpdate = 0;
        double maxScale = 0.0, minScale = 1E+20;                      // :762-807
        do {
            std::vector<double> vec, val;
            SymEigen(estCov, n, vec, val);
            bool positiveDefinite = true;
            for (int i = 0; i < n; ++i) {
                double eigen = val[i];
                if (maxScale < std::abs(eigen)) maxScale = std::abs(eigen);
                if (minScale > std::abs(eigen)) minScale = std::abs(eigen);
                if (eigen < 0) positiveDefinite = false;
            }
            if (positiveDefinite) break;
            for (int i = 0; i < n; ++i) {
                double r = estCovTrace * 1E-6;
                r /= n;
                r = std::abs(r);
                if (estCov[(size_t)i * n + i] < r) estCov[(size_t)i * n + i] = r;
                for (int j = i + 1; j < n; ++j) {
                    estCov[(size_t)i * n + j] = 0.0;
                    estCov[(size_t)j * n + i] = 0.0;
                }
            }
        } while (true);
        curCovTrace = 0.0;                                            // :813-817
        for (int i = 0; i < n; ++i) curCovTrace += std::abs(estCov[(size_t)i * n + i]);
        estCovTrace = curCovTrace;
        maxScale = std::sqrt(maxScale);                               // :820-825
        if (maxScale < 0.1) maxScale = 0.1;
        minScale = std::sqrt(minScale);
        if (minScale < 0.01) minScale = 0.01;
        orbitLength = 2.0 * 3.14 * maxScale;                          // :828
        if (meanEpsilon > 0) {                                        // :833-837
            meanEpsilon = 0.2 * maxScale;
            if (meanEpsilon > 0.5 * minScale) meanEpsilon = 0.5 * minScale;
            if (meanEpsilon < 0.05 * maxScale) meanEpsilon = 0.05 * maxScale;
        }
        if (leapFrogSteps > 0) {                                      // :839-847
            double targetLength = 0.4 * orbitLength;
            leapFrogSteps = (int)(targetLength / std::abs(meanEpsilon));
            leapFrogSteps = 2 * (leapFrogSteps / 2 + 1);
            if (leapFrogSteps > 3 * n) leapFrogSteps = 3 * n;
            if (meanEpsilon > 0) meanEpsilon = targetLength / leapFrogSteps;
        }
        estErr = estCov;                                              // :849-850
        InvertInPlace(estErr, n);
    }
    void Start(const double* x0) {                                    // :210-269
        stepCount = 0;
        proposed.assign(n, 0.0);
        proposedMomentum.assign(n, 0.0);
        accepted.assign(x0, x0 + n);
        acceptedMomentum.assign(n, 0.0);
        central.assign(n, 0.0);
        average.assign(n, 0.0);
        acceptedPotential = Potential(accepted);
        proposed = accepted;
        proposedPotential = acceptedPotential;
        meanEpsilon = 0.05;
        reversalLen = 0.0;
        targetAcceptance = 0.65;
        currentAcceptance = targetAcceptance;
        central = accepted;
        centralPotential = acceptedPotential;
        average = accepted;
        averageTrials = 0.0;
        estCov.assign((size_t)n * n, 0.0);
        exxt.assign((size_t)n * n, 0.0);
        covTrials = 0;
        for (int i = 0; i < n; ++i) estCov[(size_t)i * n + i] = 1.0;
        estErr = estCov;
        InvertInPlace(estErr, n);
        estCovTrace = n;
        stepsRemaining = 0;
        stepsSinceUpdate = 0;
    }
    int Step(int type) {                                              // :279-401
        slot = 0;
        ++stepCount;
        ProposeMomentum(proposedMomentum, acceptedMomentum);
        double initialKinetic = Kinetic(proposedMomentum);
        double lo = 0.9 * std::abs(meanEpsilon), hi = 1.1 * std::abs(meanEpsilon);
        double epsilon = lo + (hi - lo) * Rndm();
        int okLeap = LeapFrog(proposed, proposedMomentum, accepted, epsilon, std::abs(leapFrogSteps), type);
        if (okLeap < 0) return -1;
        if (leapFrogSteps > 0) {                                      // :302-323
            if (okLeap != 2) {
                if (meanEpsilon > 0 && reversalLen > meanEpsilon) {
                    double targetEpsilon = reversalLen / 8.0;
                    double deltaEpsilon = targetEpsilon - meanEpsilon;
                    if (deltaEpsilon > 0.0) meanEpsilon += 0.1 * deltaEpsilon;
                }
                if (leapFrogSteps < 50) leapFrogSteps += 1;
            } else {
                if (reversalLen < meanEpsilon) reversalLen = std::abs(leapFrogSteps * epsilon);
                else {
                    reversalLen = 0.95 * reversalLen;
                    reversalLen += 0.05 * std::abs(leapFrogSteps * epsilon);
                }
                if (leapFrogSteps > 3) leapFrogSteps -= 1;
                if (meanEpsilon > 0) meanEpsilon *= 0.99;
            }
        }
        double proposedKinetic = Kinetic(proposedMomentum);           // :326-334
        proposedPotential = Potential(proposed);
        double proposedH = proposedPotential + proposedKinetic;
        double acceptedH = acceptedPotential + initialKinetic;
        if (okLeap && std::isfinite(proposedPotential)) {             // :336-344
            UpdateCovariance();
            UpdateErrorMatrix();
        } else {
            if (meanEpsilon > 0) meanEpsilon = 0.3 * meanEpsilon;
        }
        double delta = proposedH - acceptedH;                         // :346-387
        double trial = -std::log(1.0 * Rndm());
        if (delta > trial || !std::isfinite(delta)) {
            for (int i = 0; i < n; ++i) acceptedMomentum[i] = -acceptedMomentum[i];
            currentAcceptance = (currentAcceptance * 4999.0) / 5000.0;
        } else {
            for (int i = 0; i < n; ++i) {
                accepted[i] = proposed[i];
                acceptedMomentum[i] = proposedMomentum[i];
            }
            acceptedPotential = proposedPotential;
            currentAcceptance = (currentAcceptance * 4999.0 + 1.0) / 5000.0;
        }
        if (acceptedPotential < centralPotential) {                   // :393-395
            central = accepted;
            centralPotential = acceptedPotential;
        }
        ++step;
        return 0;
    }
};
OrcHmc* HH(void* h) { return static_cast<OrcHmc*>(h); }
}  // namespace

extern "C" {

void* orc_hmc_create(int kind, int dim, int withGradient, uint64_t seed, uint32_t chain) {
    if (!((kind >= ORC_LLH_UNIT_GAUSS && kind <= ORC_LLH_ASYM) || kind == ORC_LLH_HARD) || dim < 1 ||
        (kind == ORC_LLH_HARD && dim < 2)) { gLastError = "unsupported HMC likelihood"; return 0; }
    OrcHmc* c = new OrcHmc;
    c->n = dim;
    c->like.kind = kind;
    c->like.dim = dim;
    c->withGradient = withGradient != 0;
    c->seed = seed;
    c->chain = chain;
    return c;
}
void orc_hmc_destroy(void* h) { delete HH(h); }
int orc_hmc_set_error_matrix(void* h, const double* e, int n) {
    OrcHmc* c = HH(h);
    if (c->like.kind != ORC_LLH_DUMMY || n != c->n) { gLastError = "error matrix shape"; return -1; }
    c->like.error.assign(e, e + (size_t)n * n);
    return 0;
}
int orc_hmc_set(void* h, int field, double v) {
    OrcHmc* c = HH(h);
    if (field == ORC_HMC_ALPHA) c->alpha = v;
    else if (field == ORC_HMC_MEAN_EPSILON) c->meanEpsilon = v;
    else if (field == ORC_HMC_LEAPFROG) c->leapFrogSteps = -(int)v;      // SetLeapFrog stores -i (:190)
    else { gLastError = "unknown field"; return -1; }
    return 0;
}
int orc_hmc_start(void* h, const double* x0) { HH(h)->Start(x0); return 1; }
int orc_hmc_step(void* h, int nsteps, int type, double* potential, double* x, double* epsilon, int32_t* leapfrog) {
    OrcHmc* c = HH(h);
    for (int s = 0; s < nsteps; ++s) {
        if (c->Step(type) < 0) return -1;
        if (potential) potential[s] = c->acceptedPotential;
        if (epsilon) epsilon[s] = c->meanEpsilon;
        if (leapfrog) leapfrog[s] = c->leapFrogSteps;
        if (x) std::copy(c->accepted.begin(), c->accepted.end(), x + (size_t)s * c->n);
    }
    return 0;
}
int orc_hmc_get_state(void* h, double* s, double* acc, double* mom, double* cen, double* avg, double* cov, double* err) {
    OrcHmc* c = HH(h);
    if (s) {
        s[ORC_HS_ACCEPTANCE] = c->currentAcceptance;
        s[ORC_HS_MEAN_EPSILON] = c->meanEpsilon;
        s[ORC_HS_LEAPFROG] = c->leapFrogSteps;
        s[ORC_HS_REVERSAL_LEN] = c->reversalLen;
        s[ORC_HS_ACCEPTED_POTENTIAL] = c->acceptedPotential;
        s[ORC_HS_PROPOSED_POTENTIAL] = c->proposedPotential;
        s[ORC_HS_CENTRAL_POTENTIAL] = c->centralPotential;
        s[ORC_HS_POTENTIAL_COUNT] = c->potentialCount;
        s[ORC_HS_GRADIENT_COUNT] = c->gradientCount;
        s[ORC_HS_STEP_COUNT] = c->stepCount;
        s[ORC_HS_COV_TRIALS] = c->covTrials;
        s[ORC_HS_AVERAGE_TRIALS] = c->averageTrials;
        s[ORC_HS_EST_COV_TRACE] = c->estCovTrace;
        s[ORC_HS_CUR_COV_TRACE] = c->curCovTrace;
        s[ORC_HS_ORBIT_LENGTH] = c->orbitLength;
        s[ORC_HS_STEPS_REMAINING] = c->stepsRemaining;
        s[ORC_HS_STEPS_SINCE_UPDATE] = c->stepsSinceUpdate;
    }
    if (acc) std::copy(c->accepted.begin(), c->accepted.end(), acc);
    if (mom) std::copy(c->acceptedMomentum.begin(), c->acceptedMomentum.end(), mom);
    if (cen) std::copy(c->central.begin(), c->central.end(), cen);
    if (avg) std::copy(c->average.begin(), c->average.end(), avg);
    if (cov) std::copy(c->estCov.begin(), c->estCov.end(), cov);
    if (err) std::copy(c->estErr.begin(), c->estErr.end(), err);
    return 0;
}

}  // extern "C"
