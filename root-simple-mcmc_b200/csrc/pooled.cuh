// pooled.cuh -- ensemble-pooled adaptation (BASELINE.json config 3: "adaptive
// covariance pooled across chains").  NOT in the reference, where every
// TProposeAdaptiveStep is private to its chain (TSimpleMCMC.H:547): a new
// mode, off by default, for ensembles where per-chain covariance state
// (n(n+1)/2 + n^2 doubles per chain, read and written every step) is the
// bottleneck.
//
// All chains share ONE proposal decomposition U, estimated from the pooled
// sufficient statistics of every accepted point of every chain on every GPU:
//     S = (count, sum x, sum x x^T)       [1 + n + n(n+1)/2 doubles]
//   * kPoolAccumulate adds the current accepted points of the local chains;
//   * every K steps S is all-reduced over the GPUs (NCCL, one small message);
//   * kPoolFactor turns S into mean / covariance / U = chol(covariance) with
//     the same warp Cholesky as the per-chain path, on every rank redundantly.
// Each chain keeps its own step size sigma with the reference's acceptance
// driven adaptation (TSimpleMCMC.H:1734-1776) and rescales it by the trace
// ratio when the shared covariance changes (as UpdateProposal does, :1042).
#pragma once
#include "proposal.cuh"

namespace smcmc {

struct PooledState {
    double* stats;      // accumulated S, this rank (all-reduced copy in statsAll)
    double* statsAll;   // S summed over ranks at the last exchange
    double* cov;        // packed lower triangle of the pooled covariance
    double* decomp;     // n x n row-major shared U
    double* mean;       // n
    double* trace;      // [0] trace of the pooled covariance behind `decomp`
};

// S += sum over local chains.  Block = 256 threads; thread t owns statistics
// t, t+256, ...; a block walks a slice of the chains with the points staged
// through shared memory as y = (1, x_0 .. x_{n-1}): every statistic is a product
// y_a y_b (count = y_0 y_0, sum x_i = y_{i+1} y_0, sum x_i x_j = y_{i+1} y_{j+1}),
// so the inner loop is branch-free: two LDS, one multiply, one add per statistic
// and chain.  A chain that is not running is staged as y = 0.
constexpr int kPoolOwn = 6;                 // statistics per thread and pass
constexpr int kPoolTile = 32;               // chains per staged tile
__global__ void __launch_bounds__(256)
kPoolAccumulate(const double* __restrict__ xAcc, const ChainScalars* __restrict__ sc, int chains, int n,
                double* stats) {
    extern __shared__ double tile[];          // kPoolTile x (n + 2): y, then a zero for unused slots
    const int tri = n * (n + 1) / 2;
    const int nstat = 1 + n + tri;
    const int row = n + 2;
    const int perBlock = (chains + gridDim.x - 1) / gridDim.x;
    const int first = blockIdx.x * perBlock;
    const int last = min(chains, first + perBlock);
    for (int pass = 0; pass * 256 * kPoolOwn < nstat; ++pass) {
        double acc[kPoolOwn];
        int ea[kPoolOwn], eb[kPoolOwn];
#pragma unroll
        for (int o = 0; o < kPoolOwn; ++o) {
            acc[o] = 0.0;
            const int k = (pass * kPoolOwn + o) * 256 + threadIdx.x;      // 0: count, 1..n: sum x, then packed x x^T
            ea[o] = eb[o] = n + 1;                                        // the zero slot
            if (k == 0) ea[o] = eb[o] = 0;
            else if (k <= n) { ea[o] = k; eb[o] = 0; }
            else if (k < nstat) {
                const int p = k - 1 - n;
                int i = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
                while (i * (i + 1) / 2 > p) --i;
                while ((i + 1) * (i + 2) / 2 <= p) ++i;
                ea[o] = i + 1;
                eb[o] = p - i * (i + 1) / 2 + 1;
            }
        }
        for (int c0 = first; c0 < last; c0 += kPoolTile) {
            const int nc = min(kPoolTile, last - c0);
            __syncthreads();
            for (int k = threadIdx.x; k < nc * row; k += 256) {
                const int c = k / row, i = k - c * row;
                const bool live = sc[c0 + c].started != 0;
                double v = 0.0;
                if (live && i == 0) v = 1.0;
                else if (live && i <= n) v = xAcc[(size_t)(c0 + c) * n + (i - 1)];
                tile[k] = v;
            }
            __syncthreads();
            for (int c = 0; c < nc; ++c) {
                const double* y = tile + c * row;
#pragma unroll
                for (int o = 0; o < kPoolOwn; ++o) acc[o] += y[ea[o]] * y[eb[o]];
            }
        }
#pragma unroll
        for (int o = 0; o < kPoolOwn; ++o) {
            const int k = (pass * kPoolOwn + o) * 256 + threadIdx.x;
            if (k < nstat && acc[o] != 0.0) atomicAdd(&stats[k], acc[o]);
        }
    }
}

// The same accumulation on the FP64 tensor cores.  With y = (1, x_0 .. x_{n-1}) per
// chain (y = 0 for a chain that is not running) the statistics are the lower
// triangle of S = Y^T Y, a (n+1) x (n+1) x E GEMM whose long dimension is the
// chains: 3.4e8 flop for C3 (65536 chains x 50 dims), nothing for the DMMA pipe.
// mma.sync.m8n8k4.f64 takes A(row g, k q) and B(k q, column g) from lane 4g+q:
// with k = chain and row/column = statistic BOTH fragments are y[chain q][8t + g],
// so a warp feeds the tensor cores straight from global memory -- no shared
// memory, no CTA barrier in the loop: per group of 4 chains a lane loads 7 values
// (a block of 56 statistics; 4 x 64 contiguous bytes per load instruction), the
// next group's loads are in flight while the 28 (diagonal block: lower triangle)
// or 49 DMMAs of this group issue.  Blocks of 56 statistics, block pairs
// (bi >= bj) on blockIdx.x, chain slices on blockIdx.y, the four warps of a CTA on
// quarter slices; their fragments are summed through shared memory and go to
// `stats` with one FP64 atomic per entry and CTA (packing as kPoolAccumulate).
// The count S[0][0] is a sum of ones: exact.
constexpr int kPaBlock = 56;                // statistics per block = 7 DMMA tiles

__device__ __forceinline__ void poolDmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// y[chain][y0 + 8 t + g], t = 0..6, for this lane's chain (c = group + q)
__device__ __forceinline__ void poolLoadFragment(double (&f)[7], const double* __restrict__ xAcc,
                                                 const ChainScalars* __restrict__ sc, const int* __restrict__ mask,
                                                 int n, int c, int cLast, int y0, int g) {
    // which chains take part: the running ones (sc), or -- TSimpleHMC's pooled covariance -- the
    // chains marked in `mask`
    // (the point is loaded NEXT TO the flag, not behind it: a chain outside the range reads, and
    // discards, the last row inside it)
    const bool inside = c < cLast;
    const int cc = inside ? c : cLast - 1;
    const int flag = mask ? mask[cc] : sc[cc].started;
    const double* xr = xAcc + (size_t)cc * n;
#pragma unroll
    for (int t = 0; t < 7; ++t) {
        const int y = y0 + 8 * t + g;
        f[t] = (y >= 1 && y <= n) ? xr[y - 1] : (y == 0 ? 1.0 : 0.0);
    }
    const bool live = inside && flag != 0;
#pragma unroll
    for (int t = 0; t < 7; ++t) f[t] = live ? f[t] : 0.0;
}

template <bool kDiag>
__device__ __forceinline__ void poolWarpTiles(const double* __restrict__ xAcc, const ChainScalars* __restrict__ sc,
                                              const int* __restrict__ mask, int n, int a0, int b0, int cFirst, int cLast,
                                              int lane, double* red, int warp) {
    constexpr int kTiles = kDiag ? 28 : 49;
    const int g = lane >> 2, q = lane & 3;
    double acc[kTiles][2];
#pragma unroll
    for (int t = 0; t < kTiles; ++t) acc[t][0] = acc[t][1] = 0.0;
    double af[7], bf[7], an[7], bn[7];
    if (cFirst < cLast) {
        poolLoadFragment(af, xAcc, sc, mask, n, cFirst + q, cLast, a0, g);
        if (!kDiag) poolLoadFragment(bf, xAcc, sc, mask, n, cFirst + q, cLast, b0, g);
    }
    for (int c0 = cFirst; c0 < cLast; c0 += 4) {
        if (c0 + 4 < cLast) {
            poolLoadFragment(an, xAcc, sc, mask, n, c0 + 4 + q, cLast, a0, g);
            if (!kDiag) poolLoadFragment(bn, xAcc, sc, mask, n, c0 + 4 + q, cLast, b0, g);
        }
#pragma unroll
        for (int ta = 0; ta < 7; ++ta)
#pragma unroll
            for (int tb = 0; tb < 7; ++tb) {
                if (kDiag && tb > ta) continue;
                const int idx = kDiag ? ta * (ta + 1) / 2 + tb : ta * 7 + tb;
                poolDmma(acc[idx][0], acc[idx][1], af[ta], kDiag ? af[tb] : bf[tb]);
            }
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            af[t] = an[t];
            if (!kDiag) bf[t] = bn[t];
        }
    }
    // the warps of the CTA take turns adding their fragments (same lane, same slot)
    for (int w = 0; w < 4; ++w) {
        if (warp == w) {
#pragma unroll
            for (int t = 0; t < kTiles; ++t) {
                double* r = red + (t * 32 + lane) * 2;
                if (w == 0) { r[0] = acc[t][0]; r[1] = acc[t][1]; }
                else { r[0] += acc[t][0]; r[1] += acc[t][1]; }
            }
        }
        __syncthreads();
    }
}

// kDiag: blockIdx.x = diagonal block (bi = bj); otherwise blockIdx.x enumerates the
// pairs bi > bj row by row.  Two kernels so that the diagonal one (the only one
// for n < 56) keeps its 28 accumulator tiles in registers.
template <bool kDiag>
__global__ void __launch_bounds__(128, kDiag ? 3 : 2)
kPoolAccumulateDmma(const double* __restrict__ xAcc, const ChainScalars* __restrict__ sc, int chains, int n,
                    double* stats, int chainsPerCta, const int* __restrict__ mask = nullptr) {
    __shared__ double red[(kDiag ? 28 : 49) * 64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bi = blockIdx.x, bj = blockIdx.x;
    if (!kDiag) {
        bi = 1;
        int rest = blockIdx.x;
        while (rest >= bi) { rest -= bi; ++bi; }
        bj = rest;
    }
    const int a0 = bi * kPaBlock, b0 = bj * kPaBlock;
    constexpr bool diagonal = kDiag;
    const int first = blockIdx.y * chainsPerCta;
    const int last = min(chains, first + chainsPerCta);
    const int perWarp = ((chainsPerCta + 15) / 16) * 4;                 // a multiple of 4 chains
    const int wFirst = min(last, first + warp * perWarp), wLast = min(last, wFirst + perWarp);
    poolWarpTiles<kDiag>(xAcc, sc, mask, n, a0, b0, wFirst, wLast, lane, red, warp);
    // C fragment of tile (ta, tb): row g, columns 2q and 2q+1; the lower triangle goes out
    const int tiles = diagonal ? 28 : 49;
    for (int e = tid; e < tiles * 64; e += 128) {
        const int t = e >> 6, l = (e >> 1) & 31, h = e & 1;
        int ta, tb;
        if (diagonal) {
            ta = 0;
            int r = t;
            while (r > ta) { r -= ta + 1; ++ta; }
            tb = r;
        } else {
            ta = t / 7;
            tb = t - ta * 7;
        }
        const int ya = a0 + ta * 8 + (l >> 2), yb = b0 + tb * 8 + 2 * (l & 3) + h;
        const double v = red[e];
        if (ya > n || yb > ya || v == 0.0) continue;
        int k;
        if (ya == 0) k = 0;                                     // count
        else if (yb == 0) k = ya;                               // sum x_{ya-1}
        else k = 1 + n + (ya - 1) * ya / 2 + (yb - 1);          // sum x_i x_j, j <= i
        atomicAdd(&stats[k], v);
    }
}

// ---------------------------------------------------------------------------
// The same statistics for LARGE dimensions (n >= 64): S = Y^T Y as a shared-memory tiled
// DMMA GEMM, 64 x 64 statistics per CTA, the chains as the K dimension in steps of 16 through a
// three-stage cp.async pipeline.  (kPoolAccumulateDmma above feeds the tensor cores straight from
// global memory: right for n = 50, but at n = 500 its 7-tile blocks re-read every chain row 9 times
// with 56-byte pieces -- 0.58 ms per step at 16 384 chains, 8 TFLOP/s.)  Here the statistics are
// ordered (x_0 .. x_{n-1}, 1): the x part of a chain's row starts on a 16-byte boundary when n is
// even, the column of ones is made in shared memory.  Tile pairs bi >= bj on blockIdx.x, chain
// slices on blockIdx.y; shared-memory tiles are [chain][statistic] with row stride 68, which puts
// the 16 lanes of a half warp (4 chains x 4 statistics) on 16 different bank pairs.
// ---------------------------------------------------------------------------
constexpr int kGramB = 64, kGramK = 16, kGramLd = 68, kGramStages = 3;
constexpr int kGramStageDoubles = 2 * kGramK * kGramLd;
constexpr size_t kGramSmemBytes = (size_t)kGramStages * kGramStageDoubles * sizeof(double);

template <bool VEC16>
__device__ __forceinline__ void gramLoadTile(double* dst, const double* __restrict__ x, const unsigned char* __restrict__ liveS,
                                             int n, int s0, int c0, int cFirst, int cLast, int tid) {
    constexpr int kPieces = VEC16 ? kGramK * kGramB / 2 : kGramK * kGramB;
#pragma unroll
    for (int p = 0; p < kPieces / 128; ++p) {
        const int idx = p * 128 + tid;
        const int r = VEC16 ? idx >> 5 : idx >> 6;
        const int sl = VEC16 ? (idx & 31) * 2 : (idx & 63);
        const int c = c0 + r, st = s0 + sl;
        const bool live = c < cLast && liveS[c - cFirst] != 0;
        double* d = dst + r * kGramLd + sl;
        if (VEC16) {
            if (live && st + 1 < n) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(d)),
                             "l"(x + (size_t)c * n + st) : "memory");
            } else {
                d[0] = !live ? 0.0 : (st < n ? x[(size_t)c * n + st] : (st == n ? 1.0 : 0.0));
                d[1] = !live ? 0.0 : (st + 1 < n ? x[(size_t)c * n + st + 1] : (st + 1 == n ? 1.0 : 0.0));
            }
        } else {
            if (live && st < n) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(d)),
                             "l"(x + (size_t)c * n + st) : "memory");
            } else {
                d[0] = (live && st == n) ? 1.0 : 0.0;
            }
        }
    }
}

template <bool VEC16>
__global__ void __launch_bounds__(128, 3)
kPoolGramDmma(const double* __restrict__ x, const ChainScalars* __restrict__ sc, const int* __restrict__ mask, int chains,
              int n, double* stats, int chainsPerCta) {
    extern __shared__ __align__(16) double gramSmem[];
    // which chains of this CTA's slice take part (running, or marked): read once, not per copy
    unsigned char* liveS = reinterpret_cast<unsigned char*>(gramSmem + kGramStages * kGramStageDoubles);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int g = lane >> 2, q = lane & 3;
    int bi = 0, rest = blockIdx.x;                              // block pair (bi >= bj), row by row
    while (rest > bi) { rest -= bi + 1; ++bi; }
    const int bj = rest;
    const bool diag = bi == bj;
    const int a0 = bi * kGramB, b0 = bj * kGramB;
    const int first = blockIdx.y * chainsPerCta, last = min(chains, first + chainsPerCta);
    for (int c = first + tid; c < last; c += 128) liveS[c - first] = (mask ? mask[c] != 0 : sc[c].started != 0) ? 1 : 0;
    __syncthreads();
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int steps = (last - first + kGramK - 1) / kGramK;
    auto load = [&](int st) {
        double* base = gramSmem + (st % kGramStages) * kGramStageDoubles;
        gramLoadTile<VEC16>(base, x, liveS, n, a0, first + st * kGramK, first, last, tid);
        if (!diag) gramLoadTile<VEC16>(base + kGramK * kGramLd, x, liveS, n, b0, first + st * kGramK, first, last, tid);
    };
#pragma unroll
    for (int st = 0; st < kGramStages - 1; ++st) {
        if (st < steps) load(st);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int s = 0; s < steps; ++s) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kGramStages - 2) : "memory");
        __syncthreads();
        if (s + kGramStages - 1 < steps) load(s + kGramStages - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* As = gramSmem + (s % kGramStages) * kGramStageDoubles;
        const double* Bs = diag ? As : As + kGramK * kGramLd;
#pragma unroll
        for (int kk = 0; kk < kGramK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = As[(kk + q) * kGramLd + wm + a * 8 + g];     // A(statistic g, chain q)
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bs[(kk + q) * kGramLd + wn + b * 8 + g];     // B(chain q, statistic g)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) poolDmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // C fragment of tile (a, b): row g, columns 2q and 2q+1; statistics in the order (x, 1)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int sa = a0 + wm + a * 8 + g, sb = b0 + wn + b * 8 + 2 * q + h;
                const double v = acc[a][b][h];
                if (sa > n || sb > sa || v == 0.0) continue;
                int k;
                if (sa == n) k = (sb == n) ? 0 : 1 + sb;                       // count, sum x_sb
                else k = 1 + n + sa * (sa + 1) / 2 + sb;                       // sum x_sa x_sb, sb <= sa
                atomicAdd(&stats[k], v);
            }
}

// One CTA: S -> mean, covariance, trace, U.  Keeps the previous U when the pooled covariance is
// not (yet) positive definite.  ok[0] = 1 on success.  The factorisation is the column-ordered
// U^T U = A of the per-chain path (warpCholesky: every entry the same operations in the same
// order), with one THREAD per column instead of one lane per 32 columns: at n = 500 the single
// warp took ~25 ms per exchange -- half of the pooled step at 16 384 chains -- this takes ~1 ms.
constexpr int kPoolFactorThreads = 512;
__global__ void __launch_bounds__(kPoolFactorThreads) kPoolFactorCta(PooledState ps, int n, int* ok) {
    __shared__ double pivotS;
    __shared__ int goodS;
    const int tid = threadIdx.x;
    const int tri = n * (n + 1) / 2;
    const double count = ps.statsAll[0];
    if (!(count > (double)(n + 1))) { if (tid == 0) ok[0] = 0; return; }
    for (int i = tid; i < n; i += kPoolFactorThreads) ps.mean[i] = ps.statsAll[1 + i] / count;
    __syncthreads();
    for (int p = tid; p < tri; p += kPoolFactorThreads) {
        int i = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > p) --i;
        while ((i + 1) * (i + 2) / 2 <= p) ++i;
        const int j = p - i * (i + 1) / 2;
        ps.cov[p] = ps.statsAll[1 + n + p] / count - ps.mean[i] * ps.mean[j];
    }
    if (tid == 0) goodS = 1;
    __syncthreads();
    // factor into the second half of the U buffer, publish on success
    double* u = ps.decomp + (size_t)n * n;
    for (int c = 0; c < n; ++c) {
        // entries (c, j), j >= c: v = A(c,j) - sum_{r<c} U(r,j) U(r,c), r ascending
        for (int j = c + tid; j < n; j += kPoolFactorThreads) {
            double v = ps.cov[triIndex(j, c)];
#pragma unroll 8
            for (int r = 0; r < c; ++r) v = __dsub_rn(v, __dmul_rn(u[(size_t)r * n + j], u[(size_t)r * n + c]));
            if (j == c) {
                if (v <= 0.0) goodS = 0;
                pivotS = __dsqrt_rn(v);
            }
            u[(size_t)c * n + j] = v;                      // divided by the pivot below
        }
        __syncthreads();
        if (!goodS) break;
        const double pivot = pivotS;
        for (int j = c + tid; j < n; j += kPoolFactorThreads)
            u[(size_t)c * n + j] = (j == c) ? pivot : __ddiv_rn(u[(size_t)c * n + j], pivot);
        __syncthreads();
    }
    const bool good = goodS != 0;
    if (good) {
        for (int k = tid; k < n * n; k += kPoolFactorThreads) {
            const int i = k / n, j = k - i * n;
            ps.decomp[k] = (j < i) ? 0.0 : u[k];
        }
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < n; ++i) t += ps.cov[triIndex(i, i)];
            ps.trace[0] = t;
        }
    }
    if (tid == 0) ok[0] = good ? 1 : 0;
}

// The same on one warp (kept as the reference the CTA version is tested against).
__global__ void kPoolFactor(PooledState ps, int n, int* ok) {
    const int lane = threadIdx.x;
    const int tri = n * (n + 1) / 2;
    const double count = ps.statsAll[0];
    if (!(count > (double)(n + 1))) { if (lane == 0) ok[0] = 0; return; }
    for (int i = lane; i < n; i += 32) ps.mean[i] = ps.statsAll[1 + i] / count;
    __syncwarp();
    for (int p = lane; p < tri; p += 32) {
        int i = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > p) --i;
        while ((i + 1) * (i + 2) / 2 <= p) ++i;
        const int j = p - i * (i + 1) / 2;
        ps.cov[p] = ps.statsAll[1 + n + p] / count - ps.mean[i] * ps.mean[j];
    }
    __syncwarp();
    // factor into the second half of the U buffer, publish on success
    double* work = ps.decomp + (size_t)n * n;
    const bool good = warpCholesky(ps.cov, work, n, lane);
    if (good) {
        for (int k = lane; k < n * n; k += 32) ps.decomp[k] = work[k];
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += ps.cov[triIndex(i, i)];
        if (lane == 0) ps.trace[0] = t;
    }
    if (lane == 0) ok[0] = good ? 1 : 0;
}

// The head of Step() in pooled mode: the scalar part of UpdateState
// (:1723-1776) per chain, then the draw (:709-724) with the SHARED U.
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 8)
kProposePooled(ChainArrays a, PropSettings ps, PooledState pool, int chains, uint64_t seed,
               uint32_t chainOffset, StepRef stepRef, double* zOut) {
    extern __shared__ double smemD[];
    const uint32_t step = stepRef.get();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    if (c >= chains) return;
    const int n = ps.n;
    double* cur = smemD + (size_t)warp * 3 * n;
    double* prop = cur + n;
    double* zr = prop + n;
    ChainScalars s = a.sc[c];
    if (!s.started || s.status != 0) return;
    double* xAcc = a.xAcc + (size_t)c * n;
    double* xProp = a.xProp + (size_t)c * n;
    double* last = a.lastPoint + (size_t)c * n;
    s.totalSteps += 1;
    const double value = s.accLlh;
    for (int i = lane; i < n; i += 32) cur[i] = xAcc[i];
    __syncwarp();
    s.trials += 1;
    const bool accepted = (value != s.lastValue) || (cur[0] != last[0]);
    if (accepted) s.successes += 1;
    s.acceptance = __dmul_rn(s.acceptance, s.acceptanceTrials);
    if (accepted) s.acceptance = __dadd_rn(s.acceptance, 1.0);
    s.acceptance = __ddiv_rn(s.acceptance, __dadd_rn(s.acceptanceTrials, 1.0));
    s.acceptanceTrials = fmin(ps.accWindow, __dadd_rn(s.acceptanceTrials, 1.0));
    if (s.rigidity < 500.0 && s.rigidity > 0.0) {
        double accSigma = __dmul_rn(ps.target, __dsub_rn(1.0, ps.target));
        accSigma = __dsqrt_rn(__ddiv_rn(accSigma, ps.accWindow));
        const double dist = fabs(__dsub_rn(s.acceptance, ps.target));
        if (dist < accSigma) {
            s.rigidity = __dadd_rn(s.rigidity, __ddiv_rn(__dmul_rn(0.5, s.rigidity), ps.accWindow));
            s.rigidity = fmin(200.0, s.rigidity);
        }
        if (dist > __dmul_rn(4.0, accSigma)) {
            s.rigidity = __dsub_rn(s.rigidity, __ddiv_rn(__dmul_rn(__dmul_rn(1.618, 0.5), s.rigidity), ps.accWindow));
            s.rigidity = fmax(2.0, s.rigidity);
        }
    }
    if (s.rigidity > 0 && s.rigidity < 100.0) {
        const double ex = fmin(__ddiv_rn(1.0, 500.0), __ddiv_rn(1.0, __dmul_rn(s.rigidity, ps.accWindow)));
        s.sigma = __dmul_rn(s.sigma, pow(__ddiv_rn(s.acceptance, ps.target), ex));
    }
    // the shared covariance changed since this chain last looked: keep the
    // step length in units of the new trace (UpdateProposal :1042-1043)
    const double poolTrace = pool.trace[0];
    if (poolTrace > 0.0 && poolTrace != s.sigmaTrace) {
        s.sigma = __dmul_rn(s.sigma, __dsqrt_rn(__ddiv_rn(s.sigmaTrace, poolTrace)));
        s.sigmaTrace = poolTrace;
    }
    s.lastValue = value;
    for (int i = lane; i < n; i += 32) last[i] = cur[i];

    const uint32_t gchain = chainOffset + (uint32_t)c;
    for (int pr = lane; 2 * pr < n; pr += 32) {
        double v0, v1;
        drawPair(ps, seed, gchain, step, pr, v0, v1);
        const int i = 2 * pr;
        zr[i] = ps.type[i] == 1 ? v0 : __dmul_rn(s.sigma, v0);
        if (i + 1 < n) zr[i + 1] = ps.type[i + 1] == 1 ? v1 : __dmul_rn(s.sigma, v1);
    }
    __syncwarp();
    if (zOut) {
        // TENSOR path (n >= 128): the contraction z . U of all chains runs as one
        // DMMA GEMM (contraction.cuh) and kProposePooledFinish completes the step
        for (int i = lane; i < n; i += 32) {
            const bool uniform = ps.type[i] == 1;
            zOut[(size_t)c * n + i] = uniform ? 0.0 : zr[i];
            if (uniform) xProp[i] = zr[i];
        }
        if (lane == 0) a.sc[c] = s;
        return;
    }
    const double* u = pool.decomp;
    for (int j = lane; j < n; j += 32) {
        double p;
        if (ps.type[j] == 1) p = zr[j];
        else {
            p = cur[j];
            if (!ps.anyUniform) {
#pragma unroll 8
                for (int i = 0; i <= j; ++i) p = __dadd_rn(p, __dmul_rn(zr[i], u[(size_t)i * n + j]));
            } else {
#pragma unroll 4
                for (int i = 0; i <= j; ++i) {
                    if (ps.type[i] == 1) continue;
                    p = __dadd_rn(p, __dmul_rn(zr[i], u[(size_t)i * n + j]));
                }
            }
        }
        xProp[j] = p;
        prop[j] = p;
    }
    __syncwarp();
    if (ps.stepRMSWindow > 0) {
        double sqr = 0.0;
        for (int i = 0; i < n; ++i) {
            const double d = __dsub_rn(prop[i], cur[i]);
            sqr = __dadd_rn(sqr, __dmul_rn(d, d));
        }
        double ms = __dmul_rn(s.stepRMS, s.stepRMS);
        ms = __dmul_rn(ms, (double)s.stepRMSTrials);
        ms = __dadd_rn(ms, sqr);
        ms = __ddiv_rn(ms, __dadd_rn((double)s.stepRMSTrials, 1.0));
        s.stepRMSTrials = min(ps.stepRMSWindow, s.stepRMSTrials + 1);
        s.stepRMS = __dsqrt_rn(ms);
    }
    if (lane == 0) a.sc[c] = s;
}

// TENSOR path, after the GEMM y = (sigma z) . U: x' = x + y on the Gaussian
// dimensions (uniform ones were written by kProposePooled) and the step-RMS
// tracker (TSimpleMCMC.H:391-406).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
kProposePooledFinish(ChainArrays a, PropSettings ps, int chains, const double* __restrict__ y) {
    extern __shared__ double smemD[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    if (c >= chains) return;
    const int n = ps.n;
    double* diff = smemD + (size_t)warp * n;
    ChainScalars s = a.sc[c];
    if (!s.started || s.status != 0) return;
    const double* xAcc = a.xAcc + (size_t)c * n;
    double* xProp = a.xProp + (size_t)c * n;
    for (int j = lane; j < n; j += 32) {
        const double cur = xAcc[j];
        double p;
        if (ps.type[j] == 1) p = xProp[j];
        else {
            p = __dadd_rn(cur, y[(size_t)c * n + j]);
            xProp[j] = p;
        }
        diff[j] = __dsub_rn(p, cur);
    }
    __syncwarp();
    if (ps.stepRMSWindow > 0) {
        double sqr = 0.0;
        for (int i = 0; i < n; ++i) sqr = __dadd_rn(sqr, __dmul_rn(diff[i], diff[i]));
        double ms = __dmul_rn(s.stepRMS, s.stepRMS);
        ms = __dmul_rn(ms, (double)s.stepRMSTrials);
        ms = __dadd_rn(ms, sqr);
        ms = __ddiv_rn(ms, __dadd_rn((double)s.stepRMSTrials, 1.0));
        s.stepRMSTrials = min(ps.stepRMSWindow, s.stepRMSTrials + 1);
        s.stepRMS = __dsqrt_rn(ms);
        if (lane == 0) a.sc[c] = s;
    }
}

// kProposePooled for small dimensions (the shared U fits shared memory next to a
// tile of chains): one CTA of 256 threads owns kPooledTileChains chains.  The
// scalar part of UpdateState runs one THREAD per chain, the draws one thread per
// (chain, dimension), the contraction with the shared U one thread per (chain,
// column) -- instead of one warp per chain with 31 lanes idle through the scalars.
constexpr int kPooledTileChains = 32;
constexpr int kPooledTileThreads = 256;
__host__ __device__ inline size_t pooledTileSmem(int n) {
    return ((size_t)n * n + (size_t)2 * kPooledTileChains * (n + 1) + kPooledTileChains) * sizeof(double);
}

__global__ void __launch_bounds__(kPooledTileThreads)
kProposePooledTile(ChainArrays a, PropSettings ps, PooledState pool, int chains, uint64_t seed,
                   uint32_t chainOffset, StepRef stepRef) {
    extern __shared__ double smemD[];
    const uint32_t step = stepRef.get();
    const int n = ps.n;
    const int ld = n + 1;                                    // padded rows: conflict-free column walks
    double* uS = smemD;                                      // n x n, the shared U
    double* cur = uS + (size_t)n * n;                        // [chain][n+1] accepted points
    double* zr = cur + (size_t)kPooledTileChains * ld;       // draws, then sigma * r
    double* dif = cur;                                       // squared step per dimension: entry (c, j) replaces
                                                             // x_j of chain c once its only reader has used it
    double* sig = zr + (size_t)kPooledTileChains * ld;       // [chain] fSigma, or NaN for a chain that does not run
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * kPooledTileChains;
    const int nc = min(kPooledTileChains, chains - c0);

    for (int k = tid; k < n * n; k += kPooledTileThreads) uS[k] = pool.decomp[k];
    for (int k = tid; k < nc * n; k += kPooledTileThreads) {
        const int c = k / n, i = k - c * n;
        cur[c * ld + i] = a.xAcc[(size_t)c0 * n + k];
    }
    ChainScalars s;
    bool live = false;
    if (tid < nc) {
        // ---- one thread per chain: the scalars of UpdateState, :1723-1776 ----
        const int c = c0 + tid;
        s = a.sc[c];
        live = s.started && s.status == 0;
        if (live) {
            s.totalSteps += 1;
            const double value = s.accLlh;
            const bool accepted = updateStateScalars(s, ps, value, a.xAcc[(size_t)c * n], a.lastPoint[(size_t)c * n]);
            (void)accepted;
            // the shared covariance changed since this chain last looked: keep the
            // step length in units of the new trace (UpdateProposal :1042-1043)
            const double poolTrace = pool.trace[0];
            if (poolTrace > 0.0 && poolTrace != s.sigmaTrace) {
                s.sigma = __dmul_rn(s.sigma, __dsqrt_rn(__ddiv_rn(s.sigmaTrace, poolTrace)));
                s.sigmaTrace = poolTrace;
            }
            s.lastValue = value;
        }
        sig[tid] = live ? s.sigma : nan("");
    }
    // ---- one thread per (chain, dimension): the draws, :709-719 -------------------
    // (one thread per chain and PAIR of dimensions: both normals from one Philox block)
    // (warps 1..7: warp 0 is busy with the scalars above -- pow, divisions, a square root per chain --
    // and the CTA would wait for it at the barrier if it took a share of the draws as well)
    const int npairs = (n + 1) >> 1;
    for (int k = tid - 32; k < nc * npairs; k += kPooledTileThreads - 32) {
        if (k < 0) break;
        const int c = k / npairs, pr = k - c * npairs;
        const uint32_t gchain = chainOffset + (uint32_t)(c0 + c);
        double v0, v1;
        drawPair(ps, seed, gchain, step, pr, v0, v1);
        zr[c * ld + 2 * pr] = v0;
        if (2 * pr + 1 < n) zr[c * ld + 2 * pr + 1] = v1;
    }
    __syncthreads();
    for (int k = tid; k < nc * n; k += kPooledTileThreads) {
        const int c = k / n, i = k - c * n;
        if (!(ps.anyUniform && ps.type[i] == 1)) zr[c * ld + i] = __dmul_rn(sig[c], zr[c * ld + i]);   // fSigma*r
        if (!isnan(sig[c])) a.lastPoint[(size_t)c0 * n + k] = cur[c * ld + i];                         // :1829-1830
    }
    __syncthreads();
    // ---- one thread per (chain, column): x'_j = x_j + sum_{i<=j} (fSigma r_i) U(i,j), i ascending.
    // Column-major over the tile: the 32 lanes of a warp hold the SAME column j of 32 chains,
    // so a warp runs exactly j+1 terms (with consecutive columns in a warp it waited for its
    // longest column, twice the average), U(i,j) is one broadcast read and the chains' draws
    // sit in different banks (row stride n+1).
    for (int k = tid; k < kPooledTileChains * n; k += kPooledTileThreads) {
        const int j = k / kPooledTileChains, c = k - j * kPooledTileChains;
        if (c >= nc) continue;
        const double* z = zr + c * ld;
        const double x = cur[c * ld + j];
        double p = x;
        if (ps.anyUniform && ps.type[j] == 1) {
            p = z[j];
        } else if (!ps.anyUniform) {
            const double* u = uS + j;
#pragma unroll 4
            for (int i = 0; i <= j; ++i) p = __dadd_rn(p, __dmul_rn(z[i], u[i * n]));
        } else {
            for (int i = 0; i <= j; ++i)
                if (ps.type[i] != 1) p = __dadd_rn(p, __dmul_rn(z[i], uS[i * n + j]));
        }
        if (!isnan(sig[c])) a.xProp[(size_t)(c0 + c) * n + j] = p;
        const double d = __dsub_rn(p, x);
        dif[c * ld + j] = __dmul_rn(d, d);
    }
    __syncthreads();
    if (tid < nc && live) {
        if (ps.stepRMSWindow > 0) {                                       // :391-406
            const double* q = dif + tid * ld;
            double sqr = 0.0;
            for (int i = 0; i < n; ++i) sqr = __dadd_rn(sqr, q[i]);
            double ms = __dmul_rn(s.stepRMS, s.stepRMS);
            ms = __dmul_rn(ms, (double)s.stepRMSTrials);
            ms = __dadd_rn(ms, sqr);
            ms = __ddiv_rn(ms, __dadd_rn((double)s.stepRMSTrials, 1.0));
            s.stepRMSTrials = min(ps.stepRMSWindow, s.stepRMSTrials + 1);
            s.stepRMS = __dsqrt_rn(ms);
        }
        a.sc[c0 + tid] = s;
    }
}

// ut[j][i] = u[i][j]: the K-contiguous operand of the GEMM.
__global__ void kTransposeSquare(const double* __restrict__ u, double* __restrict__ ut, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int j = idx / n, i = idx - j * n;
    ut[idx] = u[(size_t)i * n + j];
}

}  // namespace smcmc
