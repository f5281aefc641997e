// hmc.cuh -- sMCMC::TSimpleHMC (TSimpleHMC.H:119-973) for E chains at once.
//
// One HMC transition of the ensemble is a short pipeline of kernels queued by
// the host (engine.cu, hmcStepOnce):
//
//   kHmcBegin      ProposeMomentum (:554-570), initial kinetic energy (:292),
//                  the jittered step size (:297), LeapFrog's set-up (:586-611)
//   repeat k = 0 .. max_c |fLeapFrogSteps_c|:
//     gradient     PotentialGradient (:467-532) at the current positions of all
//                  chains: kDummyGradient (user gradient of TDummyLogLikelihood,
//                  a batched X.E^T contraction), the finite-difference pipeline
//                  (:417-444), kHmcCovariantGradient (:447-454) or zero
//     kHmcKickDrift  the k-th momentum update of LeapFrog (:618-648), the
//                  U-turn test (:633-638) and the next position update
//   likelihood of the proposed points (the engine's ordinary evaluate())
//   kHmcPost       leap-frog / step-size adaptation (:302-323), proposed kinetic
//                  energy (:326), UpdateCovariance (:665-695) and the trigger
//                  part of UpdateErrorMatrix (:703-719)
//   kHmcErrorMatrix (rare, only the chains that triggered) the rest of
//                  UpdateErrorMatrix (:727-858)
//   kHmcAccept     the Metropolis test on the Hamiltonian (:346-395)
//
// Chains have their own number of leap-frog steps; a chain whose trajectory is
// complete ignores the remaining iterations.  One warp owns one chain in the
// elementwise kernels (lanes across the dimension).  Parity rules as in
// proposal.cuh: the reference's operation order with __dXXX_rn intrinsics,
// sequential sums stay sequential.
//
// HBM layout, chain-major: qAcc, pAcc, qProp, pProp, p0, grad, central, average
// : double[E][n]; exxt : double[E][n(n+1)/2] packed lower triangle of fEXXT;
// estErr : double[E][n*n] (fEstimatedError, only when the covariant gradient is
// enabled); sc : HmcScalars[E].  fEstimatedCovariance is not stored: every entry
// is fEXXT(i,j) - fAveragePoint[i]*fAveragePoint[j] (:688-689) and is rebuilt
// where it is read; the diagonal that a positive-definiteness repair leaves
// behind (:792-806) is kept in repairedDiag until the next step overwrites it.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "proposal.cuh"
#include "simple_likelihoods.cuh"
#include "smcmc_b200.h"
#include "smcmc_rng.h"

namespace smcmc {

struct __align__(16) HmcScalars {
    double meanEpsilon;       // fMeanEpsilon              :902
    double reversalLen;       // fReversalLen
    double acceptance;        // fCurrentAcceptance
    double epsilon;           // the jittered step size of the step in flight :297
    double accPotential;      // fAcceptedPotential
    double propPotential;     // fProposedPotential
    double centralPotential;  // fCentralPotential
    double initialKinetic;    // :292
    double deltaH;            // proposed - accepted Hamiltonian :346
    double averageTrials;     // fAveragePointTrials
    double covTrials;         // fCovarianceTrials
    double orbitLength;       // fEstimatedOrbitLength
    double estCovTrace;       // fEstimatedCovarianceTrace
    double curCovTrace;       // fCurrentCovarianceTrace
    int leapFrogSteps;        // fLeapFrogSteps (negative = fixed by the user, :190)
    int steps;                // |fLeapFrogSteps| of the step in flight
    int okLeap;               // LeapFrog's return value
    int stepCount;            // fStepCount
    int potentialCount;       // fPotentialCount
    int gradientCount;        // fPotentialGradientCount
    int stepsRemaining;       // fStepsRemaining
    int stepsSinceUpdate;     // fStepsSinceUpdate
    int needUpdate;           // UpdateErrorMatrix passed its trigger this step
    int status;
    int started;
    int repaired;             // fEstimatedCovariance currently holds the repaired (diagonal) matrix
};
static_assert(sizeof(HmcScalars) == 160, "HmcScalars layout");

// Shared state of the pooled covariance estimate: what fCovarianceTrials, fEstimatedCovarianceTrace,
// fCurrentCovarianceTrace, fEstimatedOrbitLength, fStepsRemaining, fStepsSinceUpdate are for one
// chain (TSimpleHMC.H:665-858), kept once for the ensemble.
struct HmcPooled {
    double trials;            // samples (chain-steps) behind poolAverage / poolExxt
    double estCovTrace;
    double curCovTrace;
    double orbitLength;
    double maxScale, minScale;        // sqrt of the largest / smallest |eigenvalue| of the pooled covariance, clamped (:820-825)
    double averagePotential;          // Potential(average point) at the last update (:729)
    int stepsRemaining, stepsSinceUpdate, stepCount, needUpdate, repaired, updates;
};

struct HmcArrays {
    double* qAcc;       // fAccepted
    double* pAcc;       // fAcceptedMomentum
    double* qProp;      // fProposed
    double* pProp;      // fProposedMomentum
    double* p0;         // LeapFrog's copy of the starting momentum (:587)
    double* grad;
    double* gradCur;    // fused tensor path (kHmcLeapCached): the gradient at fAccepted, or nullptr ...
    double* gradEnd;    // ... and the one at fProposed, which replaces it when the step is accepted
    double* central;    // fCentralPoint
    double* average;    // fAveragePoint
    double* exxt;       // fEXXT, packed lower triangle
    double* exxtT;      // per chain: fCovarianceTrials before this step's UpdateCovariance, NaN = no update
    // deferred fEXXT update (deferK > 0, see kHmcExxtFlush): the points and trial counts of the
    // UpdateCovariance calls not yet applied to exxt, and the diagonal kept up to date
    double* ring;       // [E][deferK][n]
    double* ringT;      // [E][deferK]
    int* pending;       // [E] entries of the ring in use
    double* exxtDiag;   // [E][n]
    int deferK;
    double* estErr;     // fEstimatedError or nullptr
    double* repairedDiag;
    // ensemble-pooled covariance (kHmcPooled*, below): one running mean / E[x x^T] for all chains
    int pooled;         // 1: UpdateCovariance / UpdateErrorMatrix run on the pooled estimate
    int* poolMask;      // [E] chains whose UpdateCovariance call happens this step
    double* poolStats;  // [1 + n + n(n+1)/2] this step's (count, sum x, sum x x^T) over the marked chains
    double* poolExxt;   // [n(n+1)/2] pooled fEXXT
    double* poolAverage;// [n]        pooled fAveragePoint
    double* poolDiag;   // [n]        the repaired diagonal (:792-806)
    struct HmcPooled* pool;
    HmcScalars* sc;
    int* leapSteps;     // copy of sc.steps for the gradient kernels
    int* counters;      // [0] max steps of this transition, [1] chains that need UpdateErrorMatrix,
                        // [2] running chains WITHOUT a trajectory in this transition (steps < 1),
                        // [3] 2^20 - (shortest trajectory of this transition), 0 when there is none
    int* updateList;
    double* eigScratch; // slots of 2*n*n doubles
    int* eigLocks;
    int eigSlots;
};

// Sum buf[0..n) in index order; every lane returns the same value.
__device__ __forceinline__ double warpSeqSum(const double* buf, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = __dadd_rn(s, buf[i]);
    return s;
}

// Row copies of one chain by its warp: 16 bytes per lane and four loads in flight before the
// first store when the rows allow it (n even: every row starts 16-byte aligned), instead of one
// 8-byte load per lane at a time -- the pointers of HmcArrays may alias as far as the compiler
// knows, so it never hoists a load over a store by itself.  NEG: dst = -src (:364-366).
template <bool NEG>
__device__ __forceinline__ void warpCopyRow(double* dst, const double* src, int n, int lane) {
    if ((n & 1) == 0) {
        const int m = n >> 1;
        const double2* s2 = reinterpret_cast<const double2*>(src);
        double2* d2 = reinterpret_cast<double2*>(dst);
        for (int i = lane; i < m; i += 128) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + 32 * u < m) v[u] = s2[i + 32 * u];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + 32 * u < m) {
                    if (NEG) { v[u].x = -v[u].x; v[u].y = -v[u].y; }
                    d2[i + 32 * u] = v[u];
                }
        }
    } else {
        for (int i = lane; i < n; i += 32) dst[i] = NEG ? -src[i] : src[i];
    }
}

// ---------------------------------------------------------------------------
// Start (:210-269); the likelihood of the starting points is in llh[].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
kHmcStart(HmcArrays a, int n, int chains, const double* __restrict__ llh, int firstStart) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (c >= chains) return;
    const size_t row = (size_t)c * n;
    const size_t tri = (size_t)n * (n + 1) / 2;
    HmcScalars s = a.sc[c];
    s.stepCount = 0;                                          // :211
    s.potentialCount += 1;                                    // SetPosition -> Potential :204
    s.accPotential = -llh[c];
    s.propPotential = s.accPotential;                         // :225
    s.meanEpsilon = 0.05;                                     // :229
    s.reversalLen = 0.0;
    s.acceptance = 0.65;                                      // :234-235
    s.centralPotential = s.accPotential;                      // :240
    s.averageTrials = 0.0;                                    // :244
    s.covTrials = 0.0;                                        // :253
    s.estCovTrace = (double)n;                                // :263
    s.stepsRemaining = 0;                                     // :265-266
    s.stepsSinceUpdate = 0;
    s.needUpdate = 0;
    s.status = 0;
    s.started = 1;
    s.repaired = 0;
    for (int i = lane; i < n; i += 32) {
        const double x = a.qAcc[row + i];
        a.qProp[row + i] = x;                                 // :224
        a.central[row + i] = x;                               // :239
        a.average[row + i] = x;                               // :243
        if (firstStart) {                                     // resize() zero-fills only new elements
            a.pAcc[row + i] = 0.0;
            a.pProp[row + i] = 0.0;
        }
    }
    if (!a.pooled)
        for (size_t k = lane; k < tri; k += 32) a.exxt[(size_t)c * tri + k] = 0.0;    // :258
    if (a.deferK > 0) {
        for (int i = lane; i < n; i += 32) a.exxtDiag[row + i] = 0.0;
        if (lane == 0) a.pending[c] = 0;
    }
    if (a.estErr) {                                           // :261-262: the inverse of the identity
        double* e = a.estErr + (size_t)c * n * n;
        for (int k = lane; k < n * n; k += 32) e[k] = (k / n == k % n) ? 1.0 : 0.0;
    }
    if (lane == 0) a.sc[c] = s;
}

// SetPosition (:202-205).
__global__ void kHmcSetPosition(HmcArrays a, int chains, const double* __restrict__ llh) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chains) return;
    a.sc[c].potentialCount += 1;
    a.sc[c].accPotential = -llh[c];
}

// Broadcast a scalar setter to every chain (SetMeanEpsilon :181, SetLeapFrog :190).
__global__ void kHmcSetScalar(HmcScalars* sc, int chains, int field, double value) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chains) return;
    if (field == SMCMC_HMC_MEAN_EPSILON) sc[c].meanEpsilon = value;
    else if (field == SMCMC_HMC_LEAPFROG) sc[c].leapFrogSteps = -(int)value;
}

// ---------------------------------------------------------------------------
// Head of Step (:286-300) and of LeapFrog (:586-611).
// Dynamic shared memory: n doubles per warp.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
kHmcBegin(HmcArrays a, int n, int chains, double alpha, uint64_t seed, uint32_t chainOffset,
          uint32_t step) {
    extern __shared__ double smemD[];
    __shared__ double kinetic[kWarpsPerBlock];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    double* buf = smemD + (size_t)warp * n;
    HmcScalars s;
    bool run = c < chains;                                                // (the same for the whole warp)
    if (run) {
        s = a.sc[c];
        run = s.started && s.status == 0;
        if (!run && lane == 0) a.leapSteps[c] = -1;
    }
    const size_t row = (size_t)c * n;
    const uint32_t gchain = chainOffset + (uint32_t)c;
    // ProposeMomentum, :554-570.  alpha has been clamped on the host the way
    // the reference clamps its member (:558, :565).  The terms p*p/2.0 of KineticEnergy (:535-542)
    // go to shared memory as the momenta are made.
    uint32_t slot = 0;
    if (run) {
        s.stepCount += 1;                                                 // :286
        if (alpha >= 1.0) {
            for (int i = lane; i < n; i += 32) {
                const double p = __ddiv_rn(a.pAcc[row + i], alpha);
                a.pProp[row + i] = p;
                a.p0[row + i] = p;
                buf[i] = __ddiv_rn(__dmul_rn(p, p), 2.0);
            }
        } else {
            const double w = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(alpha, alpha)));
            for (int pr = lane; 2 * pr < n; pr += 32) {
                // the accepted momenta first: their load is under way while the normals are made
                const int i0 = 2 * pr;
                const bool two = i0 + 1 < n;
                const double pa0 = a.pAcc[row + i0], pa1 = two ? a.pAcc[row + i0 + 1] : 0.0;
                // slots 2 pr and 2 pr + 1: the two branches of one Box-Muller block (smcmc_rng.h)
                double z0 = 0.0, z1 = 0.0;
                smcmc_normal_pair(seed, gchain, step, (uint32_t)pr, SMCMC_STREAM_STEP, &z0, &z1);
                const double g0 = __dadd_rn(0.0, __dmul_rn(1.0, z0));     // Gaus(0,1)
                const double p0 = __dadd_rn(__dmul_rn(alpha, pa0), __dmul_rn(w, g0));
                a.pProp[row + i0] = p0;
                a.p0[row + i0] = p0;                                      // :587
                buf[i0] = __ddiv_rn(__dmul_rn(p0, p0), 2.0);
                if (two) {
                    const double g1 = __dadd_rn(0.0, __dmul_rn(1.0, z1));
                    const double p1 = __dadd_rn(__dmul_rn(alpha, pa1), __dmul_rn(w, g1));
                    a.pProp[row + i0 + 1] = p1;
                    a.p0[row + i0 + 1] = p1;
                    buf[i0 + 1] = __ddiv_rn(__dmul_rn(p1, p1), 2.0);
                }
            }
            slot = (uint32_t)n;
        }
    }
    // KineticEnergy is a sum in index order: one dependent chain of n additions per chain.  The
    // block's chains take one LANE each of warp 0 (rows n doubles apart: different banks) instead of
    // every warp walking its own chain with all 32 lanes in step -- a quarter of the instructions,
    // which were two fifths of this kernel's.
    __syncthreads();
    if (warp == 0 && lane < kWarpsPerBlock) kinetic[lane] = warpSeqSum(smemD + (size_t)lane * n, n);
    __syncthreads();
    if (!run) return;
    s.initialKinetic = kinetic[warp];                                     // :292
    {                                                                     // :297-298
        const double lo = __dmul_rn(0.9, fabs(s.meanEpsilon));
        const double hi = __dmul_rn(1.1, fabs(s.meanEpsilon));
        const double u = smcmc_uniform(seed, gchain, step, slot, SMCMC_STREAM_STEP);
        s.epsilon = __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), u));       // TRandom::Uniform(a,b)
    }
    s.steps = abs(s.leapFrogSteps);                                       // :300
    s.okLeap = 1;                                                         // :589
    if (s.steps < 1) {                                                    // :598-611
        for (int i = lane; i < n; i += 32) {
            const double p = a.pProp[row + i];
            const double mv = __ddiv_rn(__dmul_rn(s.epsilon, __dadd_rn(p, p)), 2.0);
            a.qProp[row + i] = __dadd_rn(a.qAcc[row + i], mv);
        }
    } else {
        warpCopyRow<false>(a.qProp + row, a.qAcc + row, n, lane);        // :586
    }
    if (lane == 0) {
        a.sc[c] = s;
        a.leapSteps[c] = s.steps;
        atomicMax(&a.counters[0], s.steps);
        if (s.steps < 1) atomicAdd(&a.counters[2], 1);
        else atomicMax(&a.counters[3], (1 << 20) - min(s.steps, 1 << 20));
    }
}

// ---------------------------------------------------------------------------
// After gradient number k of a trajectory (k = 0 .. steps): the momentum update
// that uses it (:618-620, :628-630, :646-648), the U-turn test (:633-638) and
// the position update that precedes gradient k+1 (:624-626, :641-643).
// countPotentials = calls of Potential() one gradient made (2n for finite
// differences, :436-438).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
kHmcKickDrift(HmcArrays a, int n, int chains, int k, int countPotentials,
              double* keepStart = nullptr, double* keepEnd = nullptr /* the gradient of k = 0 / of k = steps is kept here */) {
    extern __shared__ double smemD[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    if (c >= chains) return;
    const int steps = a.leapSteps[c];
    if (steps < 1 || k > steps) return;
    double* buf = smemD + (size_t)warp * n;
    const size_t row = (size_t)c * n;
    HmcScalars* sp = a.sc + c;
    const double eps = sp->epsilon;
    const bool half = (k == 0) || (k == steps);
    // four elements per lane are loaded before any of them is stored: the arrays of HmcArrays
    // may alias as far as the compiler knows, and a load behind a store would wait for it
    const bool drift = k < steps;
    for (int i0 = lane; i0 < n; i0 += 128) {
        double g[4], pp[4], p0v[4], q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            const bool in = i < n;
            g[u] = in ? a.grad[row + i] : 0.0;
            pp[u] = in ? a.pProp[row + i] : 0.0;
            p0v[u] = (in && !half) ? a.p0[row + i] : 0.0;
            q[u] = (in && drift) ? a.qProp[row + i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            if (i >= n) continue;
            double kick = __dmul_rn(eps, g[u]);
            if (half) kick = __ddiv_rn(kick, 2.0);
            const double p = __dsub_rn(pp[u], kick);
            a.pProp[row + i] = p;
            if (k == 0 && keepStart) keepStart[row + i] = g[u];
            if (k == steps && keepEnd) keepEnd[row + i] = g[u];
            if (!half) buf[i] = __dmul_rn(p, p0v[u]);                     // :635
            if (drift) a.qProp[row + i] = __dadd_rn(q[u], __dmul_rn(eps, p));
        }
    }
    __syncwarp();
    bool reversed = false;
    if (!half) {
        // :637-638 needs only the SIGN of the reference's ordered sum of the n products.  A
        // lane-parallel sum S and A = sum |t_i| settle it whenever |S| > 2 n u A (u = 2^-53):
        // both summation orders are within (n-1) u A of the exact sum, so they share its sign.
        // Otherwise (a knife edge, or a NaN) the ordered sum itself is formed.
        double sPar = 0.0, aPar = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double t = buf[i];
            sPar += t;
            aPar += fabs(t);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sPar += __shfl_xor_sync(0xffffffffu, sPar, o);
            aPar += __shfl_xor_sync(0xffffffffu, aPar, o);
        }
        if (fabs(sPar) > (double)n * 2.33e-16 * aPar) reversed = sPar < 0.0;
        else {
            const double inner = warpSeqSum(buf, n);
            reversed = !(inner >= 0.0);
        }
    }
    if (lane == 0) {
        sp->gradientCount += 1;                                           // :469
        if (countPotentials) sp->potentialCount += countPotentials;
        if (reversed) sp->okLeap = 2;
    }
}

// ---------------------------------------------------------------------------
// User gradient of TDummyLogLikelihood (TDummyLogLikelihood.H:34-42) with the
// sign flip of PotentialGradient (:478-487), for all chains at once:
//   g = 0;  g -= Error(i,j)*p[j]  (j ascending);  grad[i] = -g.
// A 64 x 64 x 8 shared-memory tiled contraction X[E x n] . Error^T with 4 x 4
// outputs per thread.  Multiply and subtract are separate roundings and the j
// order is the reference's, so every entry is bit-identical to the host loop.
// ---------------------------------------------------------------------------
constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 8;

__global__ void __launch_bounds__(256)
kDummyGradient(const double* __restrict__ x, const double* __restrict__ err, double* __restrict__ grad,
               const int* __restrict__ leapSteps, int k, int chains, int n) {
    __shared__ double As[kGemmBK][kGemmBM + 2];
    __shared__ double Bs[kGemmBK][kGemmBN + 2];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int c0 = blockIdx.y * kGemmBM, i0 = blockIdx.x * kGemmBN;
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.0;
    const int lr = threadIdx.x >> 2;          // 0..63: row of the tile this thread loads
    const int lk = (threadIdx.x & 3) * 2;     // 0,2,4,6
    for (int k0 = 0; k0 < n; k0 += kGemmBK) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int kk = k0 + lk + d;
            double av = 0.0, bv = 0.0;
            if (kk < n) {
                if (c0 + lr < chains) av = x[(size_t)(c0 + lr) * n + kk];
                if (i0 + lr < n) bv = err[(size_t)(i0 + lr) * n + kk];
            }
            As[lk + d][lr] = av;
            Bs[lk + d][lr] = bv;
        }
        __syncthreads();
        const int kmax = min(kGemmBK, n - k0);
        for (int kk = 0; kk < kmax; ++kk) {
            double av[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) av[r] = As[kk][ty * 4 + r];
#pragma unroll
            for (int q = 0; q < 4; ++q) bv[q] = Bs[kk][tx * 4 + q];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] = __dsub_rn(acc[r][q], __dmul_rn(bv[q], av[r]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int c = c0 + ty * 4 + r;
        if (c >= chains) continue;
        const int steps = leapSteps[c];
        if (steps < 1 || k > steps) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + tx * 4 + q;
            if (i < n) grad[(size_t)c * n + i] = -acc[r][q];
        }
    }
}

// CovariantGradient (:447-454): grad[i] = sum_j fEstimatedError(i,j)*(point[j]-fAveragePoint[j]).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
kHmcCovariantGradient(HmcArrays a, int n, int chains, int k) {
    extern __shared__ double smemD[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    if (c >= chains) return;
    const int steps = a.leapSteps[c];
    if (steps < 1 || k > steps) return;
    double* d = smemD + (size_t)warp * n;
    const size_t row = (size_t)c * n;
    for (int j = lane; j < n; j += 32) d[j] = __dsub_rn(a.qProp[row + j], a.average[row + j]);
    __syncwarp();
    const double* e = a.estErr + (size_t)c * n * n;
    for (int i = lane; i < n; i += 32) {
        double g = 0.0;
        for (int j = 0; j < n; ++j) g = __dadd_rn(g, __dmul_rn(e[(size_t)i * n + j], d[j]));
        a.grad[row + i] = g;
    }
}

// Force a flat gradient (:525-529).
__global__ void kHmcZeroGradient(HmcArrays a, int n, int chains) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (size_t)chains * n) a.grad[idx] = 0.0;
}

// FiniteDifferenceGradient (:417-444), stage 1: the 2n displaced copies of each
// chain's position.  work[((c*n + i)*2 + s)*n + j].
__global__ void kHmcFdPoints(const double* __restrict__ q, double* __restrict__ work, int chainBase,
                             int chainCount, int n) {
    const size_t total = (size_t)chainCount * n * 2 * n;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int j = (int)(idx % n);
    size_t r = idx / n;
    const int sgn = (int)(r & 1);
    r >>= 1;
    const int i = (int)(r % n);
    const int c = (int)(r / n);
    double v = q[(size_t)(chainBase + c) * n + j];
    if (j == i) {
        const double du = 0.01;                                           // :434
        v = __dsub_rn(v, du);                                             // :435
        if (sgn) v = __dadd_rn(v, __dmul_rn(2.0, du));                    // :437
    }
    work[idx] = v;
}

// ... stage 2: grad[i] = 0.5*(u2-u1)/du with u = -log(likelihood) (:439, :413).
__global__ void kHmcFdGradient(const double* __restrict__ llh, double* __restrict__ grad,
                               const int* __restrict__ leapSteps, int k, int chainBase, int chainCount, int n) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)chainCount * n) return;
    const int c = chainBase + (int)(idx / n);
    const int steps = leapSteps[c];
    if (steps < 1 || k > steps) return;
    const double u1 = -llh[idx * 2], u2 = -llh[idx * 2 + 1];
    grad[(size_t)chainBase * n + idx] = __ddiv_rn(__dmul_rn(0.5, __dsub_rn(u2, u1)), 0.01);
}

// ---------------------------------------------------------------------------
// After the trajectory and the likelihood of the proposed point: :302-344.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
kHmcPost(HmcArrays a, int n, int chains, const double* __restrict__ llhProp, double covWindow,
         int countGradients /* the fused leap-frog launches do not count: steps + 1 gradients were made (:469) */) {
    extern __shared__ double smemD[];
    __shared__ double kinetic[kWarpsPerBlock];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    double* buf = smemD + (size_t)warp * n;
    const size_t row = (size_t)c * n;
    const size_t tri = (size_t)n * (n + 1) / 2;
    // the kinetic energy of the proposed momenta (:326): terms by the chain's warp, the ordered sums
    // one lane per chain (as in kHmcBegin)
    HmcScalars s;
    bool run = c < chains;
    if (run) {
        s = a.sc[c];
        run = s.started && s.status == 0;
    }
    if (run)
        for (int i = lane; i < n; i += 32) {
            const double p = a.pProp[row + i];
            buf[i] = __ddiv_rn(__dmul_rn(p, p), 2.0);
        }
    __syncthreads();
    if (warp == 0 && lane < kWarpsPerBlock) kinetic[lane] = warpSeqSum(smemD + (size_t)lane * n, n);
    __syncthreads();
    if (!run) return;
    const double proposedKinetic = kinetic[warp];
    s.needUpdate = 0;
    if (countGradients && s.steps >= 1) s.gradientCount += s.steps + 1;
    if (lane == 0) a.exxtT[c] = __longlong_as_double(-1ll);                // NaN: no UpdateCovariance this step

    if (s.leapFrogSteps > 0) {                                            // :302-323
        if (s.okLeap != 2) {
            if (s.meanEpsilon > 0 && s.reversalLen > s.meanEpsilon) {
                const double targetEpsilon = __ddiv_rn(s.reversalLen, 8.0);
                const double deltaEpsilon = __dsub_rn(targetEpsilon, s.meanEpsilon);
                if (deltaEpsilon > 0.0) s.meanEpsilon = __dadd_rn(s.meanEpsilon, __dmul_rn(0.1, deltaEpsilon));
            }
            if (s.leapFrogSteps < 50) s.leapFrogSteps += 1;
        } else {
            const double len = fabs(__dmul_rn((double)s.leapFrogSteps, s.epsilon));
            if (s.reversalLen < s.meanEpsilon) s.reversalLen = len;
            else {
                s.reversalLen = __dmul_rn(0.95, s.reversalLen);
                s.reversalLen = __dadd_rn(s.reversalLen, __dmul_rn(0.05, len));
            }
            if (s.leapFrogSteps > 3) s.leapFrogSteps -= 1;
            if (s.meanEpsilon > 0) s.meanEpsilon = __dmul_rn(s.meanEpsilon, 0.99);
        }
    }
    s.potentialCount += 1;                                                // :327
    s.propPotential = -llhProp[c];
    const double proposedH = __dadd_rn(s.propPotential, proposedKinetic);  // :333-334
    const double acceptedH = __dadd_rn(s.accPotential, s.initialKinetic);
    s.deltaH = __dsub_rn(proposedH, acceptedH);                           // :346

    if (a.pooled) {
        // UpdateCovariance / UpdateErrorMatrix act on the estimate pooled over the ensemble: this
        // chain only says whether its call happens (:336) -- kPoolAccumulateDmma adds the marked
        // chains' accepted points, kHmcPooledFold / kHmcPooledTrigger do the rest once per step
        const bool call = s.okLeap && isfinite(s.propPotential);
        if (lane == 0) a.poolMask[c] = call ? 1 : 0;
        if (!call && s.meanEpsilon > 0) s.meanEpsilon = __dmul_rn(0.3, s.meanEpsilon);     // :343
        if (lane == 0) a.sc[c] = s;
        return;
    }
    if (s.okLeap && isfinite(s.propPotential)) {                          // :336
        // ---- UpdateCovariance, :665-695 ----------------------------------
        s.stepsSinceUpdate += 1;
        s.stepsRemaining -= 1;
        {
            const double t = s.averageTrials, t1 = __dadd_rn(t, 1.0);
            for (int i = lane; i < n; i += 32) {
                double v = __dmul_rn(a.average[row + i], t);
                v = __dadd_rn(v, a.qAcc[row + i]);
                v = __ddiv_rn(v, t1);
                a.average[row + i] = v;
            }
            s.averageTrials = fmin(covWindow, t1);
        }
        for (int i = lane; i < n; i += 32) buf[i] = a.qAcc[row + i];
        __syncwarp();
        // fEXXT itself (n(n+1)/2 entries per chain) is updated by kHmcExxtUpdate,
        // launched right after this kernel; only the diagonal is needed here.
        const double exxtT = s.covTrials, exxtT1 = __dadd_rn(exxtT, 1.0);
        if (a.deferK > 0) {
            // the update of the n(n+1)/2 entries is deferred: remember (x, T); the diagonal,
            // which the trigger below reads every step, is advanced here with the same operations
            const int slot = a.pending[c];
            double* rp = a.ring + ((size_t)c * a.deferK + slot) * n;
            for (int i = lane; i < n; i += 32) {
                const double xi = buf[i];
                rp[i] = xi;
                double d = __dmul_rn(a.exxtDiag[row + i], exxtT);
                d = __dadd_rn(d, __dmul_rn(xi, xi));
                a.exxtDiag[row + i] = __ddiv_rn(d, exxtT1);
            }
            if (lane == 0) {
                a.ringT[(size_t)c * a.deferK + slot] = exxtT;
                a.pending[c] = slot + 1;
            }
        } else if (lane == 0) a.exxtT[c] = exxtT;
        s.covTrials = fmin(covWindow, exxtT1);
        s.repaired = 0;          // fEstimatedCovariance is again fEXXT - mean mean^T
        __syncwarp();
        // ---- UpdateErrorMatrix up to its trigger, :703-719 ---------------
        if (s.leapFrogSteps != 0 && !(s.covTrials < (double)(2 * n))) {
            const double* ex = a.exxt + (size_t)c * tri;
            for (int i = lane; i < n; i += 32) {
                const double m = a.average[row + i];
                const double xi = buf[i];
                double d;
                if (a.deferK > 0) d = a.exxtDiag[row + i];                // already the updated entry
                else {
                    d = __dmul_rn(ex[triIndex(i, i)], exxtT);             // the updated diagonal entry, :683
                    d = __dadd_rn(d, __dmul_rn(xi, xi));
                    d = __ddiv_rn(d, exxtT1);
                }
                buf[i] = fabs(__dsub_rn(d, __dmul_rn(m, m)));
            }
            __syncwarp();
            s.curCovTrace = warpSeqSum(buf, n);
            const double change = fabs(__dsub_rn(s.curCovTrace, s.estCovTrace));
            bool doIt = false;
            if (s.stepsRemaining < 0) doIt = true;
            if ((double)s.stepsSinceUpdate > __dmul_rn(2.0, (double)n) &&
                change > __dmul_rn(0.01, s.estCovTrace)) doIt = true;
            if (doIt) {
                s.needUpdate = 1;
                if (lane == 0) {
                    const int slot = atomicAdd(&a.counters[1], 1);
                    a.updateList[slot] = c;
                }
            }
        }
    } else {
        if (s.meanEpsilon > 0) s.meanEpsilon = __dmul_rn(0.3, s.meanEpsilon);   // :343
    }
    if (lane == 0) a.sc[c] = s;
}

// UpdateCovariance's fEXXT(i,j) = (fEXXT(i,j) T + x_i x_j) / (T + 1), TSimpleHMC.H:678-686,
// for the chains kHmcPost marked (exxtT[c] = T): a flat streaming kernel, one CTA
// per 4096 consecutive entries of one chain's packed triangle (HBM-bound: 16 bytes
// of traffic per entry).  (i, j) is decoded once per thread and advanced with the
// packed index.
constexpr int kExxtThreads = 256;
constexpr int kExxtPerBlock = 4096;
__global__ void __launch_bounds__(kExxtThreads)
kHmcExxtUpdate(HmcArrays a, int n, int chains) {
    extern __shared__ double smemD[];
    const int c = blockIdx.y + gridDim.y * blockIdx.z;
    if (c >= chains) return;
    const double t = a.exxtT[c];
    if (!(t >= 0.0)) return;
    const double t1 = __dadd_rn(t, 1.0);
    const bool fast = t1 >= 1.0 && t1 <= 1152921504606846976.0;
    const double y = __ddiv_rn(1.0, t1);
    const long long tri = (long long)n * (n + 1) / 2;
    const long long k0 = (long long)blockIdx.x * kExxtPerBlock;
    const long long kEnd = min(tri, k0 + kExxtPerBlock);
    const double* x = a.qAcc + (size_t)c * n;
    // the rows this CTA touches end at row iMax: only x[0..iMax] is needed
    int iMax = (int)((sqrt(8.0 * (double)(kEnd - 1) + 1.0) - 1.0) * 0.5) + 1;
    if (iMax > n - 1) iMax = n - 1;
    for (int i = threadIdx.x; i <= iMax; i += kExxtThreads) smemD[i] = x[i];
    __syncthreads();
    double* ex = a.exxt + (size_t)c * tri;
    long long k = k0 + threadIdx.x;
    if (k >= kEnd) return;
    int i = (int)((sqrt(8.0 * (double)k + 1.0) - 1.0) * 0.5);
    while ((long long)i * (i + 1) / 2 > k) --i;
    while ((long long)(i + 1) * (i + 2) / 2 <= k) ++i;
    int j = (int)(k - (long long)i * (i + 1) / 2);
    constexpr int kUnroll = 4;
    for (; k < kEnd; k += (long long)kUnroll * kExxtThreads) {
        double v[kUnroll], r[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long ku = k + (long long)u * kExxtThreads;
            v[u] = ku < kEnd ? ex[ku] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            r[u] = (k + (long long)u * kExxtThreads < kEnd) ? __dmul_rn(smemD[i], smemD[j]) : 0.0;
            j += kExxtThreads;
            while (j > i) { j -= i + 1; ++i; }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long ku = k + (long long)u * kExxtThreads;
            if (ku < kEnd) {
                const double w = __dadd_rn(__dmul_rn(v[u], t), r[u]);
                ex[ku] = fast ? divideByShared(w, t1, y) : __ddiv_rn(w, t1);
            }
        }
    }
}

// Deferred form of the same update.  Rewriting every chain's n(n+1)/2 entries every
// step is the largest item of a large HMC ensemble (C4, 16384 chains x 500 dims: 32.8
// GB of traffic, 6.3 of 14.9 ms per step) although fEXXT is only READ when a chain
// passes the trigger of UpdateErrorMatrix, every ~2n steps.  With deferK > 0 kHmcPost
// only records (x, T) of each UpdateCovariance call in a per-chain ring; this kernel
// applies the recorded updates IN ORDER to each entry held in a register,
//     v <- (v T_u + x_u,i x_u,j) / (T_u + 1),  u = 0 .. pending-1,
// operation for operation what kHmcExxtUpdate does step by step (bit-identical), so
// the triangle is read and written once per deferK steps.  The host launches it for
// all chains every deferK steps and for the chains of the update list before
// kHmcErrorMatrix, and before anything else reads fEXXT.  list == nullptr: chain =
// blockIdx.y + gridDim.y * blockIdx.z; otherwise chain = list[that index].
// Dynamic shared memory: exxtFlushSmem(n, deferK).
// one recorded update of one entry: v <- (v T + x_i x_j) / (T + 1)
__device__ __forceinline__ double exxtApply(double v, double xi, double xj, double t, double t1, double y, bool fast) {
    const double r = __dmul_rn(xi, xj);
    const double w = __dadd_rn(__dmul_rn(v, t), r);
    return fast ? divideByShared(w, t1, y) : __ddiv_rn(w, t1);
}

// Shared memory: the recorded points TRANSPOSED, xsT[i][u] with a row of deferK + 2 doubles per
// dimension (the 16 values an entry needs from dimension i are contiguous: with a full ring of
// 16 they are read two at a time with LDS.128 at immediate offsets, no address arithmetic per
// update; rows of 144 bytes keep a quarter warp on different banks), then T, T + 1, 1 / (T + 1).
__host__ __device__ inline size_t exxtFlushSmem(int n, int deferK) {
    return ((size_t)n * (deferK + 2) + 3 * (size_t)deferK) * sizeof(double);
}

__global__ void __launch_bounds__(kExxtThreads)
kHmcExxtFlush(HmcArrays a, int n, int count, const int* __restrict__ list) {
    extern __shared__ __align__(16) double smemD[];
    const int idx = blockIdx.y + gridDim.y * blockIdx.z;
    if (idx >= count) return;
    const int c = list ? list[idx] : idx;
    const int p = a.pending[c];
    if (p <= 0) return;
    const int K = a.deferK;
    const int KS = K + 2;
    double* xsT = smemD;                        // [n][KS]
    double* ts = smemD + (size_t)n * KS;        // [K] T, [K] T + 1, [K] 1 / (T + 1)
    const long long tri = (long long)n * (n + 1) / 2;
    const long long k0 = (long long)blockIdx.x * kExxtPerBlock;
    const long long kEnd = min(tri, k0 + kExxtPerBlock);
    int iMax = (int)((sqrt(8.0 * (double)(kEnd - 1) + 1.0) - 1.0) * 0.5) + 1;
    if (iMax > n - 1) iMax = n - 1;
    const double* ring = a.ring + (size_t)c * K * n;
    for (int e = threadIdx.x; e < p * (iMax + 1); e += kExxtThreads) {
        const int u = e / (iMax + 1), i = e - u * (iMax + 1);
        xsT[i * KS + u] = ring[(size_t)u * n + i];
    }
    if (threadIdx.x < p) {
        const double t = a.ringT[(size_t)c * K + threadIdx.x];
        const double t1 = __dadd_rn(t, 1.0);
        ts[threadIdx.x] = t;
        ts[K + threadIdx.x] = t1;
        ts[2 * K + threadIdx.x] = __ddiv_rn(1.0, t1);
    }
    __syncthreads();
    double* ex = a.exxt + (size_t)c * tri;
    long long k = k0 + threadIdx.x;
    if (k >= kEnd) return;
    int i = (int)((sqrt(8.0 * (double)k + 1.0) - 1.0) * 0.5);
    while ((long long)i * (i + 1) / 2 > k) --i;
    while ((long long)(i + 1) * (i + 2) / 2 <= k) ++i;
    int j = (int)(k - (long long)i * (i + 1) / 2);
    // four entries per thread and pass: four independent chains of dependent updates, and the
    // three per-update scalars are read once for the four.  (Eight per thread at two CTAs per SM
    // measured slower, 51.6 against 40.2 ms per flush of C4: resident warps count for more.)
    constexpr int kUnroll = 4;
    const bool fullRing = p == K && (K == 16 || K == 8);
    for (; k < kEnd; k += (long long)kUnroll * kExxtThreads) {
        double v[kUnroll];
        int ri[kUnroll], rj[kUnroll];        // element offsets of the rows of x_i and x_j in xsT
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
            const long long kq = k + (long long)q * kExxtThreads;
            const bool in = kq < kEnd;
            v[q] = in ? ex[kq] : 0.0;
            ri[q] = (in ? i : 0) * KS;              // out of range: a staged row, the value is not stored
            rj[q] = (in ? j : 0) * KS;
            j += kExxtThreads;
            while (j > i) { j -= i + 1; ++i; }
        }
        if (fullRing) {
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
                if (u >= K) break;
                const double ta = ts[u], ta1 = ts[K + u], ya = ts[2 * K + u];
                const double tb = ts[u + 1], tb1 = ts[K + u + 1], yb = ts[2 * K + u + 1];
                const bool fa = ta1 >= 1.0 && ta1 <= 1152921504606846976.0;
                const bool fb = tb1 >= 1.0 && tb1 <= 1152921504606846976.0;
#pragma unroll
                for (int q = 0; q < kUnroll; ++q) {
                    const double2 xi = *reinterpret_cast<const double2*>(xsT + ri[q] + u);
                    const double2 xj = *reinterpret_cast<const double2*>(xsT + rj[q] + u);
                    v[q] = exxtApply(v[q], xi.x, xj.x, ta, ta1, ya, fa);
                    v[q] = exxtApply(v[q], xi.y, xj.y, tb, tb1, yb, fb);
                }
            }
        } else {
            for (int u = 0; u < p; ++u) {
                const double t = ts[u], t1 = ts[K + u], y = ts[2 * K + u];
                const bool fast = t1 >= 1.0 && t1 <= 1152921504606846976.0;
#pragma unroll
                for (int q = 0; q < kUnroll; ++q) v[q] = exxtApply(v[q], xsT[ri[q] + u], xsT[rj[q] + u], t, t1, y, fast);
            }
        }
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
            const long long kq = k + (long long)q * kExxtThreads;
            if (kq < kEnd) ex[kq] = v[q];
        }
    }
}

// after kHmcExxtFlush: the rings of the flushed chains are empty
__global__ void kHmcExxtFlushDone(HmcArrays a, int count, const int* __restrict__ list) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    a.pending[list ? list[idx] : idx] = 0;
}

// Copy the average points of the chains in the update list to a dense array.
__global__ void kHmcGatherAverage(HmcArrays a, int n, int count, double* __restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)count * n) return;
    const int c = a.updateList[idx / n];
    out[idx] = a.average[(size_t)c * n + idx % n];
}

// Eigenvalues only of the symmetric matrix in `m` (destroyed): the rotations of
// warpSymEigen without the eigenvector accumulation.  On return the diagonal
// of m holds the eigenvalues (unsorted).
__device__ void warpSymEigenValues(double* m, int n, int lane) {
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = m[(size_t)p * n + q];
                off = __dadd_rn(off, __dmul_rn(apq, apq));
            }
        if (!(off > 1e-300)) break;
        for (int p = 0; p < n; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = m[(size_t)p * n + q];
                if (apq == 0.0) continue;
                const double theta = __ddiv_rn(__dsub_rn(m[(size_t)q * n + q], m[(size_t)p * n + p]),
                                               __dmul_rn(2.0, apq));
                const double t = __ddiv_rn(theta >= 0 ? 1.0 : -1.0,
                                           __dadd_rn(fabs(theta), __dsqrt_rn(__dadd_rn(__dmul_rn(theta, theta), 1.0))));
                const double cs = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(t, t), 1.0)));
                const double sn = __dmul_rn(t, cs);
                __syncwarp();
                for (int k = lane; k < n; k += 32) {
                    const double akp = m[(size_t)k * n + p], akq = m[(size_t)k * n + q];
                    m[(size_t)k * n + p] = __dsub_rn(__dmul_rn(cs, akp), __dmul_rn(sn, akq));
                    m[(size_t)k * n + q] = __dadd_rn(__dmul_rn(sn, akp), __dmul_rn(cs, akq));
                }
                __syncwarp();
                for (int k = lane; k < n; k += 32) {
                    const double apk = m[(size_t)p * n + k], aqk = m[(size_t)q * n + k];
                    m[(size_t)p * n + k] = __dsub_rn(__dmul_rn(cs, apk), __dmul_rn(sn, aqk));
                    m[(size_t)q * n + k] = __dadd_rn(__dmul_rn(sn, apk), __dmul_rn(cs, aqk));
                }
                __syncwarp();
            }
        }
    }
}

// TMatrixD::Invert as restated by oracle/rootshim/TMatrixD.h (Gauss-Jordan with
// partial pivoting), warp cooperative: m (destroyed) and inv are n x n; lanes
// split the column index of every row operation, so each entry sees the scalar
// routine's operations in the scalar routine's order.
__device__ void warpInvert(double* m, double* inv, int n, int lane) {
    for (int k = lane; k < n * n; k += 32) inv[k] = (k / n == k % n) ? 1.0 : 0.0;
    __syncwarp();
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = fabs(m[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            const double v = fabs(m[(size_t)i * n + k]);
            if (v > best) { best = v; piv = i; }
        }
        if (piv != k) {
            for (int j = lane; j < n; j += 32) {
                double t = m[(size_t)k * n + j];
                m[(size_t)k * n + j] = m[(size_t)piv * n + j];
                m[(size_t)piv * n + j] = t;
                t = inv[(size_t)k * n + j];
                inv[(size_t)k * n + j] = inv[(size_t)piv * n + j];
                inv[(size_t)piv * n + j] = t;
            }
            __syncwarp();
        }
        const double p = m[(size_t)k * n + k];
        __syncwarp();
        for (int j = lane; j < n; j += 32) {
            m[(size_t)k * n + j] = __ddiv_rn(m[(size_t)k * n + j], p);
            inv[(size_t)k * n + j] = __ddiv_rn(inv[(size_t)k * n + j], p);
        }
        __syncwarp();
        for (int i = 0; i < n; ++i) {
            if (i == k) continue;
            const double f = m[(size_t)i * n + k];
            __syncwarp();
            if (f == 0.0) continue;
            for (int j = lane; j < n; j += 32) {
                m[(size_t)i * n + j] = __dsub_rn(m[(size_t)i * n + j], __dmul_rn(f, m[(size_t)k * n + j]));
                inv[(size_t)i * n + j] = __dsub_rn(inv[(size_t)i * n + j], __dmul_rn(f, inv[(size_t)k * n + j]));
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------
// The body of UpdateErrorMatrix (:727-858) for the chains in the update list;
// avgLlh[idx] is the likelihood at the chain's average point (:729).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
kHmcErrorMatrix(HmcArrays a, int n, int count, const double* __restrict__ avgLlh) {
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (idx >= count) return;
    const int c = a.updateList[idx];
    HmcScalars s = a.sc[c];
    const size_t row = (size_t)c * n;
    const size_t tri = (size_t)n * (n + 1) / 2;
    const double* ex = a.exxt + (size_t)c * tri;
    const double* avg = a.average + row;

    s.potentialCount += 1;                                                // :729
    const double aPot = -avgLlh[idx];
    if (aPot < s.centralPotential) {                                      // :734-738
        for (int i = lane; i < n; i += 32) a.central[row + i] = avg[i];
        s.centralPotential = aPot;
    }
    s.stepsRemaining = 2 * n + s.stepCount;                               // :758-759
    s.stepsSinceUpdate = 0;

    int slot = 0;
    if (lane == 0) {
        slot = (int)((blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) % a.eigSlots);
        while (atomicCAS(&a.eigLocks[slot], 0, 1) != 0) slot = (slot + 1) % a.eigSlots;
        __threadfence();
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    double* m = a.eigScratch + (size_t)slot * 2 * n * n;
    double* inv = m + (size_t)n * n;
    auto buildCov = [&]() {                                               // :688-689
        for (int k = lane; k < n * n; k += 32) {
            const int i = k / n, j = k - i * n;
            const double e = (j <= i) ? ex[triIndex(i, j)] : ex[triIndex(j, i)];
            m[k] = __dsub_rn(e, (j <= i) ? __dmul_rn(avg[i], avg[j]) : __dmul_rn(avg[j], avg[i]));
        }
        __syncwarp();
    };
    buildCov();
    warpSymEigenValues(m, n, lane);
    double maxScale = 0.0, minScale = 1E+20;                              // :763-781
    bool positiveDefinite = true;
    for (int i = 0; i < n; ++i) {
        const double eigen = m[(size_t)i * n + i];
        if (maxScale < fabs(eigen)) maxScale = fabs(eigen);
        if (minScale > fabs(eigen)) minScale = fabs(eigen);
        if (eigen < 0) positiveDefinite = false;
    }
    double* diag = a.repairedDiag + row;
    if (!positiveDefinite) {                                              // :792-806
        // floor the variances, drop every correlation: the matrix is diagonal
        // from here on, its eigenvalues are its diagonal, and the second pass
        // of the loop (:766-782) always ends it.
        double r = __dmul_rn(s.estCovTrace, 1E-6);
        r = __ddiv_rn(r, (double)n);
        r = fabs(r);
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            double v = __dsub_rn(ex[triIndex(i, i)], __dmul_rn(avg[i], avg[i]));
            if (v < r) v = r;
            diag[i] = v;
        }
        __syncwarp();
        __threadfence_block();
        for (int i = 0; i < n; ++i) {
            const double eigen = diag[i];
            if (maxScale < fabs(eigen)) maxScale = fabs(eigen);
            if (minScale > fabs(eigen)) minScale = fabs(eigen);
        }
        s.repaired = 1;
    }
    s.curCovTrace = 0.0;                                                  // :813-817
    for (int i = 0; i < n; ++i) {
        const double v = s.repaired ? diag[i] : __dsub_rn(ex[triIndex(i, i)], __dmul_rn(avg[i], avg[i]));
        s.curCovTrace = __dadd_rn(s.curCovTrace, fabs(v));
    }
    s.estCovTrace = s.curCovTrace;
    maxScale = __dsqrt_rn(maxScale);                                      // :820-825
    if (maxScale < 0.1) maxScale = 0.1;
    minScale = __dsqrt_rn(minScale);
    if (minScale < 0.01) minScale = 0.01;
    s.orbitLength = __dmul_rn(__dmul_rn(2.0, 3.14), maxScale);            // :828
    if (s.meanEpsilon > 0) {                                              // :833-837
        s.meanEpsilon = __dmul_rn(0.2, maxScale);
        if (s.meanEpsilon > __dmul_rn(0.5, minScale)) s.meanEpsilon = __dmul_rn(0.5, minScale);
        if (s.meanEpsilon < __dmul_rn(0.05, maxScale)) s.meanEpsilon = __dmul_rn(0.05, maxScale);
    }
    if (s.leapFrogSteps > 0) {                                            // :839-847
        const double targetLength = __dmul_rn(0.4, s.orbitLength);
        s.leapFrogSteps = (int)__ddiv_rn(targetLength, fabs(s.meanEpsilon));
        s.leapFrogSteps = 2 * (s.leapFrogSteps / 2 + 1);
        if (s.leapFrogSteps > 3 * n) s.leapFrogSteps = 3 * n;
        if (s.meanEpsilon > 0) s.meanEpsilon = __ddiv_rn(targetLength, (double)s.leapFrogSteps);
    }
    if (a.estErr) {                                                       // :849-850
        if (s.repaired) {
            for (int k = lane; k < n * n; k += 32) {
                const int i = k / n, j = k - i * n;
                m[k] = (i == j) ? diag[i] : 0.0;
            }
            __syncwarp();
        } else {
            buildCov();
        }
        warpInvert(m, inv, n, lane);
        double* e = a.estErr + (size_t)c * n * n;
        for (int k = lane; k < n * n; k += 32) e[k] = inv[k];
    }
    __syncwarp();
    __threadfence();
    if (lane == 0) {
        atomicExch(&a.eigLocks[slot], 0);
        a.sc[c] = s;
    }
}

// ===========================================================================
// Ensemble-pooled UpdateCovariance / UpdateErrorMatrix (NOT in the reference, which keeps
// one estimate per chain; SURVEY.md K11 "pooled option").  For gradient types 0 / 1 / 3 / 4 / 5
// the covariance estimate only feeds the step-size and trajectory-length tuning of
// UpdateErrorMatrix (:833-847); with thousands of chains one estimate from all of them is
// better than E private ones and costs E times less memory (n = 500, 16 384 chains: 1 MB
// instead of 16 GB) and bandwidth.  Per step: (count, sum x, sum x x^T) over the chains whose
// UpdateCovariance call happens (kPoolAccumulateDmma: Y^T Y on the FP64 tensor cores), folded
// into the running averages with the reference's update rule, weights = samples:
//     v <- (v T + sum) / (T + count),   T <- min(window, T + count).
// The trigger of UpdateErrorMatrix (:705-719) is the reference's, on the pooled quantities.
// Its body needs the largest and smallest |eigenvalue| of the covariance and whether it is
// positive definite (:762-825): a CTA-wide Cholesky factorisation decides the latter, power
// iteration on the matrix gives the largest eigenvalue and inverse iteration through the
// factor the smallest -- no full eigen-decomposition of a 500 x 500 matrix on one warp.
// ===========================================================================
constexpr int kHmcPooledMinSteps = 32;      // pooled steps before the first UpdateErrorMatrix: the chains of an
                                            // ensemble start from common points, their first samples are not a cloud

// running averages: entry k < n of the mean, then the packed second moments
__global__ void kHmcPooledFold(HmcArrays a, int n) {
    const long long tri = (long long)n * (n + 1) / 2;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n + tri) return;
    const double count = a.poolStats[0];
    if (!(count > 0.0)) return;
    const double t = a.pool->trials, t1 = __dadd_rn(t, count);
    double* dst = k < n ? a.poolAverage + k : a.poolExxt + (k - n);
    const double sum = a.poolStats[1 + k];
    *dst = __ddiv_rn(__dadd_rn(__dmul_rn(*dst, t), sum), t1);
}

// the scalar part of UpdateCovariance (:668-669, :675, :690) and UpdateErrorMatrix up to its trigger (:703-719)
__global__ void __launch_bounds__(256) kHmcPooledTrigger(HmcArrays a, int n, double window) {
    __shared__ double part[256];
    HmcPooled* p = a.pool;
    const double count = a.poolStats[0];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double m = a.poolAverage[i];
        t += fabs(__dsub_rn(a.poolExxt[triIndex(i, i)], __dmul_rn(m, m)));
    }
    part[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    p->needUpdate = 0;
    if (!(count > 0.0)) return;
    p->stepsSinceUpdate += 1;
    p->stepsRemaining -= 1;
    p->stepCount += 1;
    p->trials = fmin(window, __dadd_rn(p->trials, count));
    p->repaired = 0;
    if (p->trials < (double)(2 * n) || p->stepCount < kHmcPooledMinSteps) return;         // :705
    p->curCovTrace = part[0];
    const double change = fabs(__dsub_rn(p->curCovTrace, p->estCovTrace));
    bool doIt = false;
    if (p->stepsRemaining < 0) doIt = true;
    if ((double)p->stepsSinceUpdate > __dmul_rn(2.0, (double)n) && change > __dmul_rn(0.01, p->estCovTrace)) doIt = true;
    if (doIt) {
        p->needUpdate = 1;
        a.counters[1] = 1;
    }
}

// The body of UpdateErrorMatrix on the pooled estimate (:727-828), one CTA.  scratch: 2 n^2 + 4 n doubles.
constexpr int kHmcSpectrumThreads = 512;
__global__ void __launch_bounds__(kHmcSpectrumThreads)
kHmcPooledSpectrum(HmcArrays a, int n, const double* __restrict__ avgLlh, double* __restrict__ scratch) {
    __shared__ double red[kHmcSpectrumThreads];
    __shared__ double bcast;
    __shared__ int flag;
    const int tid = threadIdx.x;
    HmcPooled* p = a.pool;
    double* m = scratch;                         // the covariance, row-major
    double* L = m + (size_t)n * n;               // its Cholesky factor (lower triangle)
    double* x = L + (size_t)n * n;               // iteration vectors
    double* y = x + n;
    double* z = y + n;
    auto blockSum = [&](double v) {
        red[tid] = v;
        __syncthreads();
        for (int o = kHmcSpectrumThreads / 2; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        const double r = red[0];
        __syncthreads();
        return r;
    };
    for (int k = tid; k < n * n; k += kHmcSpectrumThreads) {                  // :688-689
        const int i = k / n, j = k - i * n;
        const int hi = max(i, j), lo = min(i, j);
        const double v = __dsub_rn(a.poolExxt[triIndex(hi, lo)], __dmul_rn(a.poolAverage[hi], a.poolAverage[lo]));
        m[k] = v;
        L[k] = v;
    }
    if (tid == 0) flag = 1;
    __syncthreads();
    // ---- positive definite?  right-looking Cholesky, L L^T = m
    for (int c = 0; c < n; ++c) {
        if (tid == 0) {
            const double d = L[(size_t)c * n + c];
            if (!(d > 0.0)) flag = 0;
            bcast = flag ? sqrt(d) : 1.0;
        }
        __syncthreads();
        if (!flag) break;
        const double piv = bcast;
        for (int i = c + tid; i < n; i += kHmcSpectrumThreads) L[(size_t)i * n + c] = L[(size_t)i * n + c] / piv;
        __syncthreads();
        const int rem = n - c - 1;
        for (long long k = tid; k < (long long)rem * rem; k += kHmcSpectrumThreads) {
            const int i = c + 1 + (int)(k / rem), j = c + 1 + (int)(k % rem);
            if (j <= i) L[(size_t)i * n + j] -= L[(size_t)i * n + c] * L[(size_t)j * n + c];
        }
        __syncthreads();
    }
    const bool positiveDefinite = flag != 0;
    double maxEig = 0.0, minEig = 1e20;
    if (positiveDefinite) {
        // ---- largest eigenvalue: power iteration on m (Rayleigh quotient of the last iterate)
        for (int i = tid; i < n; i += kHmcSpectrumThreads) x[i] = 1.0 + 0.37 * (double)((i * 7) % 11);
        __syncthreads();
        for (int it = 0; it < 200; ++it) {
            for (int i = tid; i < n; i += kHmcSpectrumThreads) {
                double sum = 0.0;
                const double* row = m + (size_t)i * n;
                for (int j = 0; j < n; ++j) sum += row[j] * x[j];
                y[i] = sum;
            }
            __syncthreads();
            double xx = 0.0, xy = 0.0;
            for (int i = tid; i < n; i += kHmcSpectrumThreads) {
                xx += x[i] * x[i];
                xy += x[i] * y[i];
            }
            xx = blockSum(xx);
            xy = blockSum(xy);
            const double lambda = xy / xx;
            const bool done = fabs(lambda - maxEig) <= 1e-10 * fabs(lambda);
            maxEig = lambda;
            double yy = 0.0;
            for (int i = tid; i < n; i += kHmcSpectrumThreads) yy += y[i] * y[i];
            yy = sqrt(blockSum(yy));
            for (int i = tid; i < n; i += kHmcSpectrumThreads) x[i] = y[i] / yy;
            __syncthreads();
            if (done) break;
        }
        // ---- smallest eigenvalue: inverse iteration, m^-1 x through the factor (L z = x, L^T y = z)
        for (int i = tid; i < n; i += kHmcSpectrumThreads) x[i] = 1.0 + 0.29 * (double)((i * 5) % 13);
        __syncthreads();
        double mu = 0.0;
        for (int it = 0; it < 100; ++it) {
            for (int i = tid; i < n; i += kHmcSpectrumThreads) z[i] = x[i];
            __syncthreads();
            for (int c = 0; c < n; ++c) {                                     // forward substitution, column oriented
                if (tid == 0) z[c] = z[c] / L[(size_t)c * n + c];
                __syncthreads();
                const double zc = z[c];
                for (int i = c + 1 + tid; i < n; i += kHmcSpectrumThreads) z[i] -= L[(size_t)i * n + c] * zc;
                __syncthreads();
            }
            for (int i = tid; i < n; i += kHmcSpectrumThreads) y[i] = z[i];
            __syncthreads();
            for (int c = n - 1; c >= 0; --c) {                                // back substitution with L^T
                if (tid == 0) y[c] = y[c] / L[(size_t)c * n + c];
                __syncthreads();
                const double yc = y[c];
                for (int i = tid; i < c; i += kHmcSpectrumThreads) y[i] -= L[(size_t)c * n + i] * yc;
                __syncthreads();
            }
            double xx = 0.0, xy = 0.0, yy = 0.0;
            for (int i = tid; i < n; i += kHmcSpectrumThreads) {
                xx += x[i] * x[i];
                xy += x[i] * y[i];
                yy += y[i] * y[i];
            }
            xx = blockSum(xx);
            xy = blockSum(xy);
            yy = sqrt(blockSum(yy));
            const double lambda = xy / xx;                                    // -> 1 / smallest eigenvalue
            const bool done = fabs(lambda - mu) <= 1e-8 * fabs(lambda);
            mu = lambda;
            for (int i = tid; i < n; i += kHmcSpectrumThreads) x[i] = y[i] / yy;
            __syncthreads();
            if (done) break;
        }
        minEig = 1.0 / mu;
    } else {
        // :792-806 -- floor the variances, drop every correlation: the matrix is diagonal from here on
        double r = fabs(__ddiv_rn(__dmul_rn(p->estCovTrace, 1E-6), (double)n));
        for (int i = tid; i < n; i += kHmcSpectrumThreads) {
            double v = m[(size_t)i * n + i];
            if (v < r) v = r;
            a.poolDiag[i] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double hi = 0.0, lo = 1e20;
            for (int i = 0; i < n; ++i) {
                hi = fmax(hi, fabs(a.poolDiag[i]));
                lo = fmin(lo, fabs(a.poolDiag[i]));
            }
            red[0] = hi;
            red[1] = lo;
        }
        __syncthreads();
        maxEig = red[0];
        minEig = red[1];
        __syncthreads();
    }
    double tr = 0.0;                                                          // :813-817
    for (int i = tid; i < n; i += kHmcSpectrumThreads) tr += fabs(positiveDefinite ? m[(size_t)i * n + i] : a.poolDiag[i]);
    tr = blockSum(tr);
    if (tid == 0) {
        p->repaired = positiveDefinite ? 0 : 1;
        p->curCovTrace = tr;
        p->estCovTrace = tr;
        double maxScale = sqrt(fabs(maxEig)), minScale = sqrt(fabs(minEig));   // :820-825
        if (maxScale < 0.1) maxScale = 0.1;
        if (minScale < 0.01) minScale = 0.01;
        p->maxScale = maxScale;
        p->minScale = minScale;
        p->orbitLength = 2.0 * 3.14 * maxScale;                               // :828
        p->averagePotential = -avgLlh[0];                                     // :729
        p->stepsRemaining = 2 * n + p->stepCount;                             // :758-759
        p->stepsSinceUpdate = 0;
        p->updates += 1;
    }
}

// ... and what each chain takes from it: the central point (:734-738), the step size and the
// trajectory length (:833-847).  One thread per chain.
__global__ void kHmcPooledApply(HmcArrays a, int n, int chains) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chains) return;
    HmcScalars s = a.sc[c];
    if (!s.started || s.status != 0 || s.leapFrogSteps == 0) return;          // :704
    const HmcPooled p = *a.pool;
    if (p.averagePotential < s.centralPotential) {
        for (int i = 0; i < n; ++i) a.central[(size_t)c * n + i] = a.poolAverage[i];
        s.centralPotential = p.averagePotential;
    }
    s.estCovTrace = p.estCovTrace;
    s.curCovTrace = p.curCovTrace;
    s.orbitLength = p.orbitLength;
    if (s.meanEpsilon > 0) {                                                  // :833-837
        s.meanEpsilon = __dmul_rn(0.2, p.maxScale);
        if (s.meanEpsilon > __dmul_rn(0.5, p.minScale)) s.meanEpsilon = __dmul_rn(0.5, p.minScale);
        if (s.meanEpsilon < __dmul_rn(0.05, p.maxScale)) s.meanEpsilon = __dmul_rn(0.05, p.maxScale);
    }
    if (s.leapFrogSteps > 0) {                                                // :839-847
        const double targetLength = __dmul_rn(0.4, s.orbitLength);
        s.leapFrogSteps = (int)__ddiv_rn(targetLength, fabs(s.meanEpsilon));
        s.leapFrogSteps = 2 * (s.leapFrogSteps / 2 + 1);
        if (s.leapFrogSteps > 3 * n) s.leapFrogSteps = 3 * n;
        if (s.meanEpsilon > 0) s.meanEpsilon = __ddiv_rn(targetLength, (double)s.leapFrogSteps);
    }
    a.sc[c] = s;
}

struct HmcTraceDev {
    double* potential;    // fAcceptedPotential after the step ("LogLikelihood" branch :139)
    double* points;       // fAccepted                          ("Accepted" :140)
    double* meanEpsilon;  // fMeanEpsilon                       ("MeanEpsilon" :145)
    int32_t* leapfrog;    // fLeapFrogSteps                     ("Leapfrog" :147)
    int32_t* accepted;    // 1 when the proposed point was taken
};

// The Metropolis test on the Hamiltonian and the commit, :346-395.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
kHmcAccept(HmcArrays a, int n, int chains, uint64_t seed, uint32_t chainOffset, uint32_t step,
           uint32_t slot, HmcTraceDev tr, int traceStep) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (c >= chains) return;
    HmcScalars s = a.sc[c];
    if (!s.started || s.status != 0) return;
    const size_t row = (size_t)c * n;
    const double u = __dmul_rn(1.0, smcmc_uniform(seed, chainOffset + (uint32_t)c, step, slot, SMCMC_STREAM_STEP));
    const double trial = -log(u);                                         // :347
    const bool reject = (s.deltaH > trial) || !isfinite(s.deltaH);        // :348
    if (reject) {
        warpCopyRow<true>(a.pAcc + row, a.pAcc + row, n, lane);          // :364-366
        s.acceptance = __ddiv_rn(__dmul_rn(s.acceptance, 4999.0), 5000.0);               // :367
    } else {
        warpCopyRow<false>(a.qAcc + row, a.qProp + row, n, lane);        // :380-383
        warpCopyRow<false>(a.pAcc + row, a.pProp + row, n, lane);
        if (a.gradCur) warpCopyRow<false>(a.gradCur + row, a.gradEnd + row, n, lane);
        s.accPotential = s.propPotential;                                 // :384
        s.acceptance = __ddiv_rn(__dadd_rn(__dmul_rn(s.acceptance, 4999.0), 1.0), 5000.0);   // :386
    }
    __syncwarp();
    if (s.accPotential < s.centralPotential) {                            // :393-395
        warpCopyRow<false>(a.central + row, a.qAcc + row, n, lane);
        s.centralPotential = s.accPotential;
    }
    if (lane == 0) a.sc[c] = s;
    if (traceStep >= 0) {
        const size_t r = (size_t)traceStep * chains + c;
        if (lane == 0) {
            if (tr.potential) tr.potential[r] = s.accPotential;
            if (tr.meanEpsilon) tr.meanEpsilon[r] = s.meanEpsilon;
            if (tr.leapfrog) tr.leapfrog[r] = s.leapFrogSteps;
            if (tr.accepted) tr.accepted[r] = reject ? 0 : 1;
        }
        if (tr.points) warpCopyRow<false>(tr.points + r * n, a.qAcc + row, n, lane);
    }
}

}  // namespace smcmc
