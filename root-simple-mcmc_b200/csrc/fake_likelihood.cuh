// fake_likelihood.cuh -- the event-loop likelihood of example/FakeLikelihood.H
// for many parameter points (chains) at once.
//
// Reference: FakeLikelihood::operator() (example/FakeLikelihood.H:47-81),
// FillHistograms (:188-216), SystematicCorrection::{InvariantMass,Separation,
// EventWeight,CorrectEvent} (example/SystematicCorrection.H:35-128).
//
// How the work is cut:
//  * Everything that depends only on the EVENT (3 logs, one division) is
//    computed once when the sample is uploaded: PreparedEvent, 32 bytes.
//  * Everything that depends only on the CHAIN (5 exp, erf, 2 atan) is computed
//    once per likelihood evaluation: FakeChainParams.
//  * What is left per (chain, event) pair is one exp, 7 FP64 operations in the
//    reference's order, and a bin lookup.  The corrected mass itself is never
//    exponentiated: TH1's bin of exp(l) is found by comparing l with the
//    pre-image of the bin edges under the host's exp (gEdges), which is the
//    same decision as the reference's exp-then-FindBin.
//  * An event's weight takes one of four values per chain (signal/background x
//    decay tag), so the histograms are accumulated as exact INTEGER counts per
//    weight class and turned into the reference's sequentially-rounded bin
//    contents afterwards (seqsum.h).  Integer counts make the result
//    independent of event order, of the tiling and of the number of GPUs.
//  * Pair kernel: one CTA = 256 chains (one per thread) x one chunk of events
//    of a single weight class.  Event tiles are staged into shared memory by
//    TMA bulk copies (cp.async.bulk + mbarrier, double buffered) and every
//    thread reads them by broadcast; each thread owns a private column of
//    shared-memory counters, so no atomics and no bank conflicts.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "seqsum.h"
#include "smcmc_b200.h"

namespace smcmc {

// Per-event constants (example/SystematicCorrection.H:53-62).
struct __align__(32) PreparedEvent {
    double logSigma;   // (log m - log M0) / (log(M0+S0) - log M0)
    double dLog;       // log m - log M0
    double nomLog;     // log M0
    double sep;        // evt.Separation
};
static_assert(sizeof(PreparedEvent) == 32, "PreparedEvent is 32 bytes");

// Per-chain constants.
struct FakeChainParams {
    double skewc;      // 0.3*erf(p[kMassSkew]/10)              :68
    double width;      // exp(p[kMassWidth]/10)                 :65
    double scale;      // p[kMassScale]/10                      :64
    double sepScale[2];// exp((0+p[5|6])/10)  signal, background :42-45
    double weight[4];  // [signal, signal&tag, background, background&tag] :81-117
};

// Weight classes and the layout of the count table.
//   class 0: signal, MuDk==0     -> slots [  0,100)  Close bins, Separated bins
//   class 1: signal, MuDk>0      -> slots [100,150)  DecayTag bins
//   class 2: background, MuDk==0 -> slots [150,250)
//   class 3: background, MuDk>0  -> slots [250,300)
//   class 4: data-typed (Type<0) -> slots [300,450)  Close, Separated, DecayTag
// counts[slot][point] (point fastest) are uint32.
constexpr int kFakeClasses = 4;
constexpr int kFakeSlots = 450;
__host__ __device__ constexpr int fakeClassSlotBase(int cls) {
    return cls == 0 ? 0 : cls == 1 ? 100 : cls == 2 ? 150 : cls == 3 ? 250 : 300;
}
constexpr int kIrregularClass = 5;   // events the fast path cannot take

constexpr int kPairThreads = 256;    // chains per CTA
constexpr int kPairTile = 128;       // events per shared-memory tile
constexpr int kPairChunk = 8192;     // events per CTA work item
constexpr int kPairCounterRows = 100;

// Pre-images of the 50 bin edges under the host's exp, and of the cut at 500:
//   bin (1-based) of exp(l) is 1 + #{k in 1..49 : l >= gEdges[k]},
//   the event is dropped when !(l < gEdges[50]).
__constant__ double gEdges[52];
// tan(M_PI*(trueFakes-0.5)), tan(M_PI*(trueEfficiency-0.5)), M_PI: host values.
__constant__ double gFakeConst[4];

// ---------------------------------------------------------------------------
// Upload-time preparation.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int classifyEvent(const smcmc_event& e, PreparedEvent& p) {
    double nomLog = log(e.TrueMass);                                   // :57
    double nomLogSigma = log(__dadd_rn(e.TrueMass, e.TrueMassSigma));  // :58
    nomLogSigma = __dsub_rn(nomLogSigma, nomLog);                      // :59
    double logMass = log(e.Mass);                                      // :61
    double d = __dsub_rn(logMass, nomLog);
    p.dLog = d;
    p.logSigma = __ddiv_rn(d, nomLogSigma);                            // :62
    p.nomLog = nomLog;
    p.sep = e.Separation;
    bool regular = e.Type >= 0 && isfinite(p.dLog) && isfinite(p.logSigma) &&
                   isfinite(p.nomLog) && isfinite(p.sep) && p.sep >= 0.0;
    if (!regular) return kIrregularClass;
    return (e.Type == 0 ? 0 : 2) + (e.MuDk > 0 ? 1 : 0);
}

__global__ void kFakeCountClasses(const smcmc_event* __restrict__ ev, int64_t n,
                                  unsigned long long* classCount, int forceGeneric) {
    __shared__ unsigned int local[8];
    if (threadIdx.x < 8) local[threadIdx.x] = 0;
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        PreparedEvent p;
        int cls = forceGeneric ? kIrregularClass : classifyEvent(ev[i], p);
        atomicAdd(&local[cls], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 8 && local[threadIdx.x])
        atomicAdd(&classCount[threadIdx.x], (unsigned long long)local[threadIdx.x]);
}

// Scatter events into their class segment.  Order inside a segment is
// arbitrary (integer counting does not depend on it).
__global__ void kFakeScatter(const smcmc_event* __restrict__ ev, int64_t n,
                             PreparedEvent* prepared, const int64_t* classBase,
                             unsigned long long* cursor, smcmc_event* irregular,
                             int forceGeneric) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PreparedEvent p;
    smcmc_event e = ev[i];
    int cls = forceGeneric ? kIrregularClass : classifyEvent(e, p);
    unsigned long long pos = atomicAdd(&cursor[cls], 1ull);
    if (cls == kIrregularClass) irregular[pos] = e;
    else prepared[classBase[cls] + (int64_t)pos] = p;
}

// ---------------------------------------------------------------------------
// Per-evaluation chain constants.
// ---------------------------------------------------------------------------
__global__ void kFakePrepareChains(const double* __restrict__ x, int m, int dim,
                                   double exposure, FakeChainParams* out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const double* p = x + (size_t)c * dim;
    FakeChainParams cp;
    cp.scale = __ddiv_rn(p[2], 10.0);                                  // :64
    cp.width = exp(__ddiv_rn(p[3], 10.0));                             // :65
    cp.skewc = __dmul_rn(0.3, erf(__ddiv_rn(p[4], 10.0)));             // :68
    cp.sepScale[0] = exp(__ddiv_rn(__dadd_rn(0.0, p[5]), 10.0));       // :41-45
    cp.sepScale[1] = exp(__ddiv_rn(__dadd_rn(0.0, p[6]), 10.0));
    const double pi = gFakeConst[2];
    double wSig = __dmul_rn(1.0, exp(__ddiv_rn(p[0], 10.0)));          // :88
    double wBkg = __dmul_rn(1.0, exp(__ddiv_rn(p[1], 10.0)));          // :89
    const double trueFakes = 0.05;                                     // :94-97
    double fakes = __dadd_rn(gFakeConst[0], p[7]);
    fakes = __dadd_rn(__ddiv_rn(atan(fakes), pi), 0.5);
    const double trueEff = 0.5;                                        // :106-109
    double eff = __dadd_rn(gFakeConst[1], p[8]);
    eff = __dadd_rn(__ddiv_rn(atan(eff), pi), 0.5);
    double sigTag = __dmul_rn(wSig, __ddiv_rn(fakes, trueFakes));                                   // :100
    double sigNo = __dmul_rn(wSig, __ddiv_rn(__dsub_rn(1.0, fakes), __dsub_rn(1.0, trueFakes)));    // :101
    double bkgTag = __dmul_rn(wBkg, __ddiv_rn(eff, trueEff));                                       // :112
    double bkgNo = __dmul_rn(wBkg, __ddiv_rn(__dsub_rn(1.0, eff), __dsub_rn(1.0, trueEff)));        // :113
    cp.weight[0] = __dmul_rn(sigNo, exposure);                         // :116
    cp.weight[1] = __dmul_rn(sigTag, exposure);
    cp.weight[2] = __dmul_rn(bkgNo, exposure);
    cp.weight[3] = __dmul_rn(bkgTag, exposure);
    out[c] = cp;
}

// ---------------------------------------------------------------------------
// Bin lookup.
// ---------------------------------------------------------------------------
// Exact: 0-based bin of a corrected log-mass, or 50 when the event is cut
// (FakeLikelihood.H:203-204 and the TH1 overflow bin).  NaN is cut.
__device__ __noinline__ int exactBin(double lm, int guess) {
    if (!(lm < gEdges[50])) return 50;
    int k = min(max(guess, 0), 49);
    while (k > 0 && lm < gEdges[k]) --k;
    while (k < 49 && lm >= gEdges[k + 1]) ++k;
    return k;
}

// double -> float by bit manipulation (truncation), without touching the
// FP64 pipe's conversion unit.  Only used for the bin guess.
__device__ __forceinline__ float truncToFloat(double v) {
    const unsigned hi = (unsigned)__double2hiint(v);
    const unsigned lo = (unsigned)__double2loint(v);
    const int e = (int)((hi >> 20) & 0x7ff) - 1023;
    unsigned bits;
    if (e < -60) bits = 0u;                         // |v| tiny: exp(v) == 1 to float precision
    else if (e > 7) bits = 0x43800000u;             // |v| >= 256 (also inf/NaN): saturate
    else bits = ((unsigned)(e + 127) << 23) | ((hi & 0xfffffu) << 3) | (lo >> 29);
    return __uint_as_float(bits | (hi & 0x80000000u));
}

// The common case costs FP32/SFU/integer work only: an approximate mass gives
// a candidate bin, accepted when it is farther than the approximation error
// from a bin edge.  Everything else goes through exactBin.
__device__ __forceinline__ int fastBin(double lm) {
    float mf = __expf(truncToFloat(lm));            // relative error < 2e-6
    float q = mf * 0.1f;                            // bins are 10 wide
    float qc = fminf(q, 60.0f);
    // nearest integer to qc-0.5 by the 1.5*2^23 trick (FP32 pipe, no F2I)
    float shifted = (qc - 0.5f) + 12582912.0f;
    float kf = shifted - 12582912.0f;
    float frac = qc - kf;
    int k = __float_as_int(shifted) - 0x4b400000;
    bool sure = (frac > 4e-4f) && (frac < 1.0f - 4e-4f) && (q < 49.9f) && (k >= 0);
    bool surelyOut = (q > 50.1f) && (q < 1e30f);
    if (sure) return k;
    if (surelyOut) return 50;
    return exactBin(lm, k);
}

// ---------------------------------------------------------------------------
// TMA bulk copy + mbarrier helpers (sm_90+ PTX; SASS: UBLKCP / SYNCS).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smemAddr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmaLoad1D(void* dstSmem, const void* srcGlobal, uint32_t bytes,
                                          uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smemAddr(dstSmem)),
        "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
        : "memory");
}

// ---------------------------------------------------------------------------
// The pair kernel.
// ---------------------------------------------------------------------------
struct PairLaunch {
    const PreparedEvent* events;     // all classes, class segments contiguous
    int64_t classBase[kFakeClasses]; // first event of each class
    int64_t classCount[kFakeClasses];
    int chunkBase[kFakeClasses + 1]; // prefix sum of chunks per class
    const FakeChainParams* chains;
    int numPoints;                   // chains (parameter points) to evaluate
    int pointStride;                 // row length of the count table
    uint32_t* counts;                // [kFakeSlots][pointStride]
};

template <bool TAGGED>
__device__ __forceinline__ void pairOne(const PreparedEvent& ev, double skewc, double width,
                                        double scale, double sepScale, uint32_t* myCounters) {
    // example/SystematicCorrection.H:69-74, in the reference's order
    double skew = exp(__dmul_rn(ev.logSigma, skewc));
    double lm = __dadd_rn(ev.nomLog, __dmul_rn(ev.dLog, skew));
    lm = __dadd_rn(ev.nomLog, __dmul_rn(__dsub_rn(lm, ev.nomLog), width));
    lm = __dadd_rn(lm, scale);
    int bin = fastBin(lm);                        // mass = exp(lm); TH1 bin of mass
    if (bin >= 50) return;                        // FakeLikelihood.H:203-204 / overflow
    int row = bin;
    if (!TAGGED) {
        double sep = __dmul_rn(ev.sep, sepScale); // SystematicCorrection.H:45-47
        if (!(sep < 100.0)) row += 50;            // FakeLikelihood.H:210-214
    }
    myCounters[row * kPairThreads] += 1;
}

template <bool TAGGED>
__device__ __forceinline__ void pairChunk(const PairLaunch& L, int cls, int64_t first, int count,
                                          int pointBase, PreparedEvent (*tiles)[kPairTile],
                                          uint64_t* bars, uint32_t* counters) {
    const int tid = threadIdx.x;
    const int point = pointBase + tid;
    const bool live = point < L.numPoints;
    FakeChainParams cp;
    if (live) cp = L.chains[point];
    else { cp.skewc = 0; cp.width = 1; cp.scale = 0; cp.sepScale[0] = cp.sepScale[1] = 1; }
    const double sepScale = cp.sepScale[cls >> 1];
    constexpr int rows = TAGGED ? 50 : 100;
    for (int r = 0; r < rows; ++r) counters[r * kPairThreads + tid] = 0;
    uint32_t* mine = counters + tid;

    const PreparedEvent* src = L.events + L.classBase[cls] + first;
    const int numTiles = (count + kPairTile - 1) / kPairTile;
    if (tid == 0) {
        int len = min(kPairTile, count);
        mbarExpectTx(&bars[0], (uint32_t)len * sizeof(PreparedEvent));
        tmaLoad1D(tiles[0], src, (uint32_t)len * sizeof(PreparedEvent), &bars[0]);
    }
    for (int t = 0; t < numTiles; ++t) {
        const int buf = t & 1;
        if (tid == 0 && t + 1 < numTiles) {
            // buffer buf^1 was released by the __syncthreads at the end of tile t-1
            int len = min(kPairTile, count - (t + 1) * kPairTile);
            mbarExpectTx(&bars[buf ^ 1], (uint32_t)len * sizeof(PreparedEvent));
            tmaLoad1D(tiles[buf ^ 1], src + (size_t)(t + 1) * kPairTile,
                      (uint32_t)len * sizeof(PreparedEvent), &bars[buf ^ 1]);
        }
        mbarWait(&bars[buf], (uint32_t)(t >> 1) & 1u);
        const int len = min(kPairTile, count - t * kPairTile);
        const PreparedEvent* tile = tiles[buf];
        int e = 0;
        for (; e + 1 < len; e += 2) {
            PreparedEvent a = tile[e], b = tile[e + 1];
            pairOne<TAGGED>(a, cp.skewc, cp.width, cp.scale, sepScale, mine);
            pairOne<TAGGED>(b, cp.skewc, cp.width, cp.scale, sepScale, mine);
        }
        if (e < len) pairOne<TAGGED>(tile[e], cp.skewc, cp.width, cp.scale, sepScale, mine);
        __syncthreads();
    }
    if (live) {
        const int slotBase = fakeClassSlotBase(cls);
        for (int r = 0; r < rows; ++r) {
            uint32_t v = counters[r * kPairThreads + tid];
            if (v) atomicAdd(&L.counts[(size_t)(slotBase + r) * L.pointStride + point], v);
        }
    }
}

// grid.x = (#chunks over all classes) * (#point tiles); consecutive CTAs take
// the point tiles of the same chunk, so a chunk is fetched from HBM once and
// re-read from L2 by the other tiles.
__global__ void __launch_bounds__(kPairThreads, 2)
kFakePairs(const __grid_constant__ PairLaunch L) {
    extern __shared__ __align__(128) unsigned char smemRaw[];
    PreparedEvent(*tiles)[kPairTile] = reinterpret_cast<PreparedEvent(*)[kPairTile]>(smemRaw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smemRaw + 2 * kPairTile * sizeof(PreparedEvent));
    uint32_t* counters = reinterpret_cast<uint32_t*>(smemRaw + 2 * kPairTile * sizeof(PreparedEvent) + 64);

    const int pointTiles = (L.numPoints + kPairThreads - 1) / kPairThreads;
    const int chunk = blockIdx.x / pointTiles;
    const int pointBase = (blockIdx.x - chunk * pointTiles) * kPairThreads;
    int cls = 0;
    while (cls + 1 < kFakeClasses && chunk >= L.chunkBase[cls + 1]) ++cls;
    const int64_t first = (int64_t)(chunk - L.chunkBase[cls]) * kPairChunk;
    const int count = (int)min((int64_t)kPairChunk, L.classCount[cls] - first);

    if (threadIdx.x == 0) {
        mbarInit(&bars[0], 1);
        mbarInit(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (cls & 1) pairChunk<true>(L, cls, first, count, pointBase, tiles, bars, counters);
    else pairChunk<false>(L, cls, first, count, pointBase, tiles, bars, counters);
}

constexpr size_t kPairSmemBytes =
    2 * kPairTile * sizeof(PreparedEvent) + 64 + (size_t)kPairCounterRows * kPairThreads * sizeof(uint32_t);

// Straight transcription of the per-event formula for the events the fast
// path cannot take (data-typed, negative separation, non-finite fields):
// one thread per (point, event).
__global__ void kFakePairsGeneric(const smcmc_event* __restrict__ ev, int64_t nev,
                                  const double* __restrict__ x, int m, int dim,
                                  uint32_t* counts, int pointStride) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nev * (int64_t)m) return;
    int point = (int)(idx % m);
    smcmc_event e = ev[idx / m];
    const double* p = x + (size_t)point * dim;
    double mass = e.Mass, sep = e.Separation;
    if (e.Type >= 0) {
        double nomLog = log(e.TrueMass);
        double nomLogSigma = __dsub_rn(log(__dadd_rn(e.TrueMass, e.TrueMassSigma)), nomLog);
        double logMass = log(mass);
        double logSigma = __ddiv_rn(__dsub_rn(logMass, nomLog), nomLogSigma);
        double scale = __ddiv_rn(p[2], 10.0);
        double width = exp(__ddiv_rn(p[3], 10.0));
        double skew = __dmul_rn(0.3, erf(__ddiv_rn(p[4], 10.0)));
        skew = exp(__dmul_rn(logSigma, skew));
        logMass = __dadd_rn(nomLog, __dmul_rn(__dsub_rn(logMass, nomLog), skew));
        logMass = __dadd_rn(nomLog, __dmul_rn(__dsub_rn(logMass, nomLog), width));
        logMass = __dadd_rn(logMass, scale);
        double sc = exp(__ddiv_rn(__dadd_rn(0.0, e.Type == 0 ? p[5] : p[6]), 10.0));
        sep = __dmul_rn(sep, sc);
        if (sep < 0.0) return;
        int bin = exactBin(logMass, 25);
        if (bin >= 50) return;
        int cls = (e.Type == 0 ? 0 : 2) + (e.MuDk > 0 ? 1 : 0);
        int row = bin;
        if (!(cls & 1) && !(sep < 100.0)) row += 50;
        atomicAdd(&counts[(size_t)(fakeClassSlotBase(cls) + row) * pointStride + point], 1u);
    } else {
        if (mass > 500.0 || mass < 0.0 || sep < 0.0) return;
        if (!(mass < 500.0)) return;                 // NaN or ==500: overflow bin
        int bin = (int)(50 * (mass - 0.0) / (500.0 - 0.0));
        int h = e.MuDk > 0 ? 2 : (sep < 100.0 ? 0 : 1);
        atomicAdd(&counts[(size_t)(300 + h * 50 + bin) * pointStride + point], 1u);
    }
}

// ---------------------------------------------------------------------------
// Counts -> bin contents -> log-likelihood (FakeLikelihood.H:51-80).
// Block = 32 points (lane = point) x kFinishWarps warps striding the 150 bins.
// ---------------------------------------------------------------------------
constexpr int kFinishWarps = 10;

__global__ void __launch_bounds__(32 * kFinishWarps)
kFakeFinish(const uint32_t* __restrict__ counts, int pointStride, int m,
            const FakeChainParams* __restrict__ chains, const double* __restrict__ data150,
            double* llhOut, double* histOut) {
    __shared__ double term[150][33];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int point = blockIdx.x * 32 + lane;
    const bool live = point < m;
    double w[4] = {0, 0, 0, 0};
    if (live) {
        for (int k = 0; k < 4; ++k) w[k] = chains[point].weight[k];
    }
    for (int hb = warp; hb < 150; hb += kFinishWarps) {
        const int h = hb / 50, b = hb - h * 50;
        double v = 0.0;
        if (live) {
            uint32_t nSig, nBkg, nDat;
            double wSig, wBkg;
            if (h < 2) {
                nSig = counts[(size_t)(0 + h * 50 + b) * pointStride + point];
                nBkg = counts[(size_t)(150 + h * 50 + b) * pointStride + point];
                wSig = w[0];
                wBkg = w[2];
            } else {
                nSig = counts[(size_t)(100 + b) * pointStride + point];
                nBkg = counts[(size_t)(250 + b) * pointStride + point];
                wSig = w[1];
                wBkg = w[3];
            }
            nDat = counts[(size_t)(300 + hb) * pointStride + point];
            // TH1::Fill in event order: the signal block, then the background
            // block (Simulated::MakeSample, example/Simulated.H:17-29).
            double mc = smcmc_seq_add(0.0, wSig, nSig);
            mc = smcmc_seq_add(mc, wBkg, nBkg);
            mc = smcmc_seq_add(mc, 1.0, nDat);
            if (histOut) histOut[(size_t)point * 150 + hb] = mc;
            const double d = data150[hb];
            if (mc < 0.001) mc = 0.001;                                 // :56
            v = __dsub_rn(d, mc);                                       // :57
            if (d > 0.0) v = __dadd_rn(v, __dmul_rn(d, log(__ddiv_rn(mc, d))));   // :58
        }
        term[hb][lane] = v;
    }
    __syncthreads();
    if (warp == 0 && live && llhOut) {
        double s = 0.0;                                                 // :51,59 in bin order
        for (int hb = 0; hb < 150; ++hb) s = __dadd_rn(s, term[hb][lane]);
        llhOut[point] = s;
    }
}

}  // namespace smcmc
