// unbinned_likelihood.cuh -- the unbinned mixture likelihood of BASELINE.json
// configs[4] ("ensemble sweep: chains x events of an unbinned likelihood"), for
// many chains at once.
//
// The reference has NO unbinned likelihood: this functor is
// defined here (include/smcmc_b200.h, SMCMC_LLH_UNBINNED) on top of the
// reference's per-event corrections, and its CPU checker is the builder-written
// oracle/smcmc_oracle.cc::EvalUnbinned -- parity is NOT pinned by the reference.
//
//   L(theta) = sum_e log( w_s(theta, tag_e) phi_s(m'_e) + w_b(theta, tag_e) phi_b(m'_e) )
//   log m'_e = corrected log-mass, SystematicCorrection::InvariantMass
//              (example/SystematicCorrection.H:50-79): nl + d exp(ls c) width + scale
//   w_s, w_b = the signal / background event weights of EventWeight (:81-117)
//   phi_s    = log-normal density around 135 (sigma of the log = log 1.3)
//   phi_b    = exponential density, tau = 500
//
// Unlike the binned likelihood, nothing here is a discrete decision: every
// (chain, event) pair needs three exponentials and one log1p in FP64, and the
// kernel is bound by the FP64 pipe.  One CTA = 128 chains (one per thread) x one
// chunk of events of one tag class; PreparedEvent tiles (128 x 32 B) are staged
// by TMA bulk copies, double buffered, and read by broadcast; each thread keeps
// one running sum per chunk, written to partial[chunk][chain]; kUnbinnedFinish
// adds the chunks in chunk order, so a result does not depend on scheduling.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "fake_likelihood.cuh"

namespace smcmc {

struct UnbinnedChain {
    double skewc;      // 0.3*erf(p[4]/10)
    double width;      // exp(p[3]/10)
    double scale;      // p[2]/10
    double la[2];      // log w_s + log normalisation of phi_s: untagged, tagged
    double lb[2];      // log w_b - log tau
};

constexpr int kUnbThreads = 128;
constexpr int kUnbTile = 128;
constexpr int kUnbMaxChunk = 32768;

__global__ void kUnbinnedPrepareChains(const double* __restrict__ x, int m, int dim, UnbinnedChain* out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const double* p = x + (size_t)c * dim;
    UnbinnedChain u;
    u.scale = __ddiv_rn(p[2], 10.0);
    u.width = exp(__ddiv_rn(p[3], 10.0));
    u.skewc = __dmul_rn(0.3, erf(__ddiv_rn(p[4], 10.0)));
    const double pi = gFakeConst[2];
    double fakes = __dadd_rn(gFakeConst[0], p[7]);
    fakes = __dadd_rn(__ddiv_rn(atan(fakes), pi), 0.5);
    double eff = __dadd_rn(gFakeConst[1], p[8]);
    eff = __dadd_rn(__ddiv_rn(atan(eff), pi), 0.5);
    const double wSig = __dmul_rn(1.0, exp(__ddiv_rn(p[0], 10.0)));
    const double wBkg = __dmul_rn(1.0, exp(__ddiv_rn(p[1], 10.0)));
    const double sig = log(1.3), tau = 500.0;
    const double ca = -log(__dmul_rn(sig, sqrt(__dmul_rn(2.0, pi)))), cb = -log(tau);
    u.la[0] = __dadd_rn(log(__dmul_rn(wSig, __ddiv_rn(__dsub_rn(1.0, fakes), __dsub_rn(1.0, 0.05)))), ca);
    u.la[1] = __dadd_rn(log(__dmul_rn(wSig, __ddiv_rn(fakes, 0.05))), ca);
    u.lb[0] = __dadd_rn(log(__dmul_rn(wBkg, __ddiv_rn(__dsub_rn(1.0, eff), __dsub_rn(1.0, 0.5)))), cb);
    u.lb[1] = __dadd_rn(log(__dmul_rn(wBkg, __ddiv_rn(eff, 0.5))), cb);
    out[c] = u;
}

// Upload: key = (tag << 32) | original index, so that the radix sort leaves the
// events of a tag class in their original order (the sum is reproducible).
__global__ void kUnbinnedSortKeys(const smcmc_event* __restrict__ ev, int64_t n, unsigned long long* keys,
                                  unsigned int* index, unsigned long long* tagged) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned tag = ev[i].MuDk > 0 ? 1u : 0u;
    keys[i] = ((unsigned long long)tag << 32) | (unsigned long long)i;
    index[i] = (unsigned int)i;
    if (tag) atomicAdd(tagged, 1ull);
}

__global__ void kUnbinnedGather(const smcmc_event* __restrict__ ev, int64_t n, const unsigned int* __restrict__ index,
                                PreparedEvent* prepared) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const smcmc_event e = ev[index[i]];
    PreparedEvent p;
    const double nomLog = log(e.TrueMass);
    const double nomLogSigma = __dsub_rn(log(__dadd_rn(e.TrueMass, e.TrueMassSigma)), nomLog);
    p.dLog = __dsub_rn(log(e.Mass), nomLog);
    p.logSigma = __ddiv_rn(p.dLog, nomLogSigma);
    p.nomLog = nomLog;
    p.sep = e.Separation;
    prepared[i] = p;
}

struct UnbinnedLaunch {
    const PreparedEvent* events;   // untagged events, then tagged events
    int64_t classBase[2];
    int64_t classCount[2];
    int chunkBase[3];              // prefix sum of chunks per class
    int chunkEvents;               // events per chunk, a multiple of kUnbTile
    const UnbinnedChain* chains;
    int numPoints;
    int stride;                    // row length of partial[][]
    double* partial;               // [chunks][stride]
};

__global__ void __launch_bounds__(kUnbThreads)
kUnbinnedPairs(const __grid_constant__ UnbinnedLaunch L) {
    __shared__ __align__(128) PreparedEvent tiles[2][kUnbTile];
    __shared__ __align__(8) uint64_t bars[2];
    const int tid = threadIdx.x;
    const int pointTiles = (L.numPoints + kUnbThreads - 1) / kUnbThreads;
    const int chunk = blockIdx.x / pointTiles;
    const int point = (blockIdx.x - chunk * pointTiles) * kUnbThreads + tid;
    const int cls = chunk >= L.chunkBase[1] ? 1 : 0;
    const int64_t first = (int64_t)(chunk - L.chunkBase[cls]) * L.chunkEvents;
    const int count = (int)min((int64_t)L.chunkEvents, L.classCount[cls] - first);
    const bool live = point < L.numPoints;
    UnbinnedChain cp;
    if (live) cp = L.chains[point];
    else { cp.skewc = 0.0; cp.width = 1.0; cp.scale = 0.0; cp.la[0] = cp.la[1] = cp.lb[0] = cp.lb[1] = 0.0; }
    const double la = cp.la[cls], lb = cp.lb[cls];
    const double mu = log(135.0), invSig = 1.0 / log(1.3), invTau = 1.0 / 500.0;

    if (tid == 0) {
        mbarInit(&bars[0], 1);
        mbarInit(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const PreparedEvent* src = L.events + L.classBase[cls] + first;
    const int numTiles = (count + kUnbTile - 1) / kUnbTile;
    if (tid == 0) {
        const int len = min(kUnbTile, count);
        mbarExpectTx(&bars[0], (uint32_t)len * sizeof(PreparedEvent));
        tmaLoad1D(tiles[0], src, (uint32_t)len * sizeof(PreparedEvent), &bars[0]);
    }
    double acc = 0.0;
    for (int t = 0; t < numTiles; ++t) {
        const int buf = t & 1;
        if (tid == 0 && t + 1 < numTiles) {
            const int len = min(kUnbTile, count - (t + 1) * kUnbTile);
            mbarExpectTx(&bars[buf ^ 1], (uint32_t)len * sizeof(PreparedEvent));
            tmaLoad1D(tiles[buf ^ 1], src + (size_t)(t + 1) * kUnbTile, (uint32_t)len * sizeof(PreparedEvent), &bars[buf ^ 1]);
        }
        mbarWait(&bars[buf], (uint32_t)(t >> 1) & 1u);
        const int len = min(kUnbTile, count - t * kUnbTile);
#pragma unroll 2
        for (int e = 0; e < len; ++e) {
            const PreparedEvent ev = tiles[buf][e];
            const double skew = exp(__dmul_rn(ev.logSigma, cp.skewc));
            double lm = __dadd_rn(ev.nomLog, __dmul_rn(ev.dLog, skew));
            lm = __dadd_rn(ev.nomLog, __dmul_rn(__dsub_rn(lm, ev.nomLog), cp.width));
            lm = __dadd_rn(lm, cp.scale);
            const double z = (lm - mu) * invSig;
            const double a = la - 0.5 * z * z - lm;
            const double b = lb - exp(lm) * invTau;
            const double hi = fmax(a, b), lo = fmin(a, b);
            // NaN-propagating on purpose: an event with a non-positive mass makes the likelihood NaN
            acc += hi + log1p(exp(lo - hi));
        }
        __syncthreads();
    }
    if (live) L.partial[(size_t)chunk * L.stride + point] = acc;
}

__global__ void kUnbinnedFinish(const double* __restrict__ partial, int chunks, int stride, int m, double* __restrict__ llh) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    double s = 0.0;
    for (int k = 0; k < chunks; ++k) s += partial[(size_t)k * stride + c];
    llh[c] = s;
}

}  // namespace smcmc
