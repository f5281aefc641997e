// contraction.cuh -- the dense contraction Y = X . Error^T for E chains on the
// FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64), used by the TENSOR mode of the
// TDummyLogLikelihood functors:
//   gradient of the potential (TDummyLogLikelihood.H:34-42 with the sign flip of
//   TSimpleHMC.H:478-487):   grad[c][i] =  sum_j Error(i,j) x[c][j]  = Y[c][i]
//   likelihood (TDummyLogLikelihood.H:21-31, Error symmetric):
//                            L[c] = -1/2 sum_i x[c][i] Y[c][i]
// The default EXACT mode (hmc.cuh kDummyGradient, simple_likelihoods.cuh
// kDummyLikelihood) keeps the reference's operation order and is bit-identical
// to it; this mode re-associates the sums (fused multiply-adds in the tensor
// core's order) and agrees with the reference to ~n * 2^-53 relative to
// sum |terms| -- inside the 1e-12 of the specification, not bit for bit.
//
// Tiling: CTA = 128 threads = 4 warps (2 x 2), CTA tile 64 chains x 64 outputs,
// warp tile 32 x 32 = 4 x 4 DMMA tiles (8 x 8 x 4), K in steps of 16 staged in
// shared memory through cp.async, double buffered.  Both operands are
// K-contiguous in global memory (X row = one chain, Error row = one output).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace smcmc {

constexpr int kDmmaBM = 64, kDmmaBN = 64, kDmmaBK = 16;
constexpr int kDmmaPad = 4;                      // row stride 20 doubles: the 16 lanes of a half warp
                                                 // (rows g..g+3, columns q) hit 16 different bank pairs

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// One 64 x 16 tile of a K-contiguous matrix (rows x n) into shared memory
// [64][16 + pad]; rows past `rows` and columns past `n` are zero.  Rows start on
// 16-byte boundaries when n is even (VEC16: 16-byte cp.async, 4 per thread);
// otherwise 8-byte copies.
template <bool VEC16>
__device__ __forceinline__ void dmmaLoadTile(double (*dst)[kDmmaBK + kDmmaPad], const double* __restrict__ src,
                                             int row0, int rows, int k0, int n, int tid) {
    if (VEC16) {
        // 64 rows x 8 pairs = 512 16-byte pieces, 128 threads: 4 each
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int idx = p * 128 + tid;               // 0..511
            const int r = idx >> 3, k = (idx & 7) * 2;
            double* d = &dst[r][k];
            if (row0 + r < rows && k0 + k + 1 < n) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(d)),
                             "l"(src + (size_t)(row0 + r) * n + k0 + k) : "memory");
            } else {
                d[0] = (row0 + r < rows && k0 + k < n) ? src[(size_t)(row0 + r) * n + k0 + k] : 0.0;
                d[1] = 0.0;
            }
        }
    } else {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int idx = p * 128 + tid;               // 0..1023
            const int r = idx >> 4, k = idx & 15;
            double* d = &dst[r][k];
            if (row0 + r < rows && k0 + k < n) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(d)),
                             "l"(src + (size_t)(row0 + r) * n + k0 + k) : "memory");
            } else {
                *d = 0.0;
            }
        }
    }
}

// The same tile with its rows GATHERED: tile row r is row rowMap[r] of the matrix (shared memory, 64
// entries; -1: no row) -- the rows of a tile need not be neighbours (kHmcLeapDmma, chains ordered by
// trajectory length).  A function of its own: with the row map as an optional argument of the loader
// above the compiler no longer kept the contiguous rows' addresses in registers across the K loop
// (156 -> 132 registers and a slower kernel for everybody).
template <bool VEC16>
__device__ __forceinline__ void dmmaLoadTileRows(double (*dst)[kDmmaBK + kDmmaPad], const double* __restrict__ src,
                                                 const int* rowMap, int k0, int n, int tid) {
    if (VEC16) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int idx = p * 128 + tid;
            const int r = idx >> 3, k = (idx & 7) * 2;
            double* d = &dst[r][k];
            const int row = rowMap[r];
            if (row >= 0 && k0 + k + 1 < n) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(d)),
                             "l"(src + (size_t)row * n + k0 + k) : "memory");
            } else {
                d[0] = (row >= 0 && k0 + k < n) ? src[(size_t)row * n + k0 + k] : 0.0;
                d[1] = 0.0;
            }
        }
    } else {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int idx = p * 128 + tid;
            const int r = idx >> 4, k = idx & 15;
            double* d = &dst[r][k];
            const int row = rowMap[r];
            if (row >= 0 && k0 + k < n) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(d)),
                             "l"(src + (size_t)row * n + k0 + k) : "memory");
            } else {
                *d = 0.0;
            }
        }
    }
}

// The K loop of the 64 x 64 CTA tile: kDmmaStages-deep cp.async pipeline over K steps of 16,
// 4 x 4 DMMA tiles per warp.  smem: [stages][2][64][20] doubles (dynamic).  acc is the C
// fragment of the warp tile: tile (a, b) holds row g, columns 2q and 2q+1.
constexpr int kDmmaStages = 3;
constexpr int kDmmaStageDoubles = 2 * kDmmaBM * (kDmmaBK + kDmmaPad);
constexpr size_t kDmmaSmemBytes = (size_t)kDmmaStages * kDmmaStageDoubles * sizeof(double);

template <bool VEC16, bool GATHER = false>
__device__ __forceinline__ void dmmaMainloop(double (&acc)[4][4][2], double* smem, const double* __restrict__ x,
                                             const double* __restrict__ err, int c0, int chains, int i0, int n, int tid,
                                             const int* rowMap = nullptr) {
    typedef double (*Tile)[kDmmaBK + kDmmaPad];
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int steps = (n + kDmmaBK - 1) / kDmmaBK;
#pragma unroll
    for (int st = 0; st < kDmmaStages - 1; ++st) {
        if (st < steps) {
            if (GATHER) dmmaLoadTileRows<VEC16>((Tile)(smem + st * kDmmaStageDoubles), x, rowMap, st * kDmmaBK, n, tid);
            else dmmaLoadTile<VEC16>((Tile)(smem + st * kDmmaStageDoubles), x, c0, chains, st * kDmmaBK, n, tid);
            dmmaLoadTile<VEC16>((Tile)(smem + st * kDmmaStageDoubles + kDmmaBM * (kDmmaBK + kDmmaPad)), err, i0, n, st * kDmmaBK, n, tid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int s = 0; s < steps; ++s) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kDmmaStages - 2) : "memory");
        __syncthreads();                                   // stage s landed; everybody is done with stage s-1
        const int nxt = s + kDmmaStages - 1;
        if (nxt < steps) {
            double* base = smem + (nxt % kDmmaStages) * kDmmaStageDoubles;
            if (GATHER) dmmaLoadTileRows<VEC16>((Tile)base, x, rowMap, nxt * kDmmaBK, n, tid);
            else dmmaLoadTile<VEC16>((Tile)base, x, c0, chains, nxt * kDmmaBK, n, tid);
            dmmaLoadTile<VEC16>((Tile)(base + kDmmaBM * (kDmmaBK + kDmmaPad)), err, i0, n, nxt * kDmmaBK, n, tid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        Tile As = (Tile)(smem + (s % kDmmaStages) * kDmmaStageDoubles);
        Tile Bs = (Tile)(smem + (s % kDmmaStages) * kDmmaStageDoubles + kDmmaBM * (kDmmaBK + kDmmaPad));
#pragma unroll
        for (int kk = 0; kk < kDmmaBK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = As[wm + a * 8 + g][kk + q];      // A(row g, col q)
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bs[wn + b * 8 + g][kk + q];      // B(k q, col g) = err[i][k]
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
}

// mode 0: y[c][i] = sum_j err[i][j] x[c][j]                      (gradient of the potential)
// mode 1: partial[c][blockIdx.x] = sum_{i in this block} x[c][i] * Y[c][i]   (for the likelihood)
// leapSteps / k: as kDummyGradient (chains whose trajectory is complete are not written); may be null.
// Dynamic shared memory: kDmmaSmemBytes.
template <bool VEC16>
__global__ void __launch_bounds__(128, 3)
kDummyContractDmma(const double* __restrict__ x, const double* __restrict__ err, double* __restrict__ out,
                   const int* __restrict__ leapSteps, int k, int chains, int n, int mode) {
    extern __shared__ __align__(16) double dmmaSmem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;     // warp tile origin inside the CTA tile
    const int c0 = blockIdx.y * kDmmaBM, i0 = blockIdx.x * kDmmaBN;
    const int g = lane >> 2, q = lane & 3;                     // fragment coordinates
    double acc[4][4][2];
    dmmaMainloop<VEC16>(acc, dmmaSmem, x, err, c0, chains, i0, n, tid);
    // C fragment: row g, columns 2q and 2q+1 of each 8 x 8 tile
    if (mode == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int c = c0 + wm + a * 8 + g;
            if (c >= chains) continue;
            if (leapSteps) {
                const int st = leapSteps[c];
                if (st < 1 || k > st) continue;
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = i0 + wn + b * 8 + 2 * q;
                if (i < n) out[(size_t)c * n + i] = acc[a][b][0];
                if (i + 1 < n) out[(size_t)c * n + i + 1] = acc[a][b][1];
            }
        }
    } else {
        // x[c][i] * Y[c][i] summed over this block's 64 outputs: per chain row, over b, the two columns, the
        // four lanes of a quad (shuffle), and the two warps that share the rows (shared memory)
        __shared__ double part[kDmmaBM][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int c = c0 + wm + a * 8 + g;
            double sum = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = i0 + wn + b * 8 + 2 * q;
                if (c < chains && i < n) sum = fma(x[(size_t)c * n + i], acc[a][b][0], sum);
                if (c < chains && i + 1 < n) sum = fma(x[(size_t)c * n + i + 1], acc[a][b][1], sum);
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (q == 0) part[wm + a * 8 + g][warp & 1] = sum;
        }
        __syncthreads();
        if (tid < kDmmaBM && c0 + tid < chains) out[(size_t)(c0 + tid) * gridDim.x + blockIdx.x] = part[tid][0] + part[tid][1];
    }
}

// Host side: launch with the copy width the operands allow.
inline void launchDummyContractDmma(cudaStream_t stream, const double* x, const double* err, double* out,
                                    const int* leapSteps, int k, int chains, int n, int mode) {
    static bool ready = false;
    if (!ready) {
        cudaFuncSetAttribute(kDummyContractDmma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDmmaSmemBytes);
        cudaFuncSetAttribute(kDummyContractDmma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDmmaSmemBytes);
        ready = true;
    }
    dim3 grid((n + kDmmaBN - 1) / kDmmaBN, (chains + kDmmaBM - 1) / kDmmaBM);
    if ((n & 1) == 0)
        kDummyContractDmma<true><<<grid, 128, kDmmaSmemBytes, stream>>>(x, err, out, leapSteps, k, chains, n, mode);
    else
        kDummyContractDmma<false><<<grid, 128, kDmmaSmemBytes, stream>>>(x, err, out, leapSteps, k, chains, n, mode);
}

// ---------------------------------------------------------------------------
// One leap-frog stage of TSimpleHMC::LeapFrog (TSimpleHMC.H:612-648) for the dense
// Gaussian, FUSED with its gradient: the CTA that holds the 64 x 64 tile of
// grad = X . Error^T in its accumulators applies, element by element,
//     p  <- p - eps g          (/ 2 for the first and the last gradient, :618-620, :646-648)
//     q' <- q + eps p          (when another gradient follows, :624-626, :641-643)
// and the per-chain partial sums of p . p0 over its 64 dimensions for the U-turn test
// (:633-638).  Nothing of the gradient goes to HBM and the kick / drift pass over four
// E x n arrays (kHmcKickDrift) disappears.  q is double buffered (qIn -> qOut): other CTAs
// of the same launch still read qIn as their GEMM operand.  Chains that do not take part
// in gradient k (trajectory complete) copy q through, so both buffers stay current.
// The U-turn partials of gradient k are summed (fixed order over the column blocks) by the
// blockIdx.x = 0 CTAs of the NEXT launch, which exists for every chain that needs it
// (the test is made for k = 1 .. steps-1).  TENSOR mode is specified to 1e-12 of the
// reference, not bit for bit, so the sign of that sum is taken as it is.
// ---------------------------------------------------------------------------
struct LeapFused {
    const double* qIn;
    double* qOut;
    double* p;               // fProposedMomentum, updated in place
    const double* p0;        // the momentum the trajectory started with
    const double* epsilon;   // per chain: &sc[c].epsilon, stride `scalarStride` doubles
    int* okLeap;             // per chain: &sc[c].okLeap, stride `scalarStride` doubles
    int scalarStride;        // sizeof(HmcScalars) / 8
    const int* leapSteps;
    double* uturn;           // [2][chains][blocks]: sum of p p0 over each column block, gradients k (parity k & 1)
    int blocks;              // column blocks = gridDim.x
    double* endPartial;      // [chains][blocks] or null: x . (Error x) over each column block at the END point of a chain's
                             // trajectory (its last gradient, k == steps) -- the potential there without another GEMM
    double* gradStart;       // null, or [chains][n]: the gradient at the START point (k = 0) is kept ...
    double* gradEnd;         // null, or [chains][n]: ... and the one at the END point (k == steps): the next step starts at one
                             // of the two, and its first gradient is the one kept (kHmcLeapCached)
    const int* order;        // null, or the chains ordered by trajectory length, longest first: row r of the launch is
                             // chain order[r]
    int gemmTiles;           // with `order`: the first gemmTiles row tiles (blockIdx.y) hold every chain that still takes
                             // part in gradient k; row tiles behind them (launched for k = 0 only) copy qIn to qOut
};

template <bool VEC16, bool ORDERED>
__global__ void __launch_bounds__(128, 3)
kHmcLeapDmma(const double* __restrict__ err, LeapFused f, int k, int chains, int n) {
    extern __shared__ __align__(16) double dmmaSmem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int c0 = blockIdx.y * kDmmaBM, i0 = blockIdx.x * kDmmaBN;
    const int g = lane >> 2, q = lane & 3;
    // ORDERED -- ragged trajectory lengths: the launch's rows are the chains in order of length
    // (f.order), so the chains that still run fill the first row tiles, and only those are launched:
    // the unordered launch runs the whole GEMM for rows that ignore its result.  A chain's last gradient
    // is taken at its end point and copies it to the second position buffer (no drift), so both buffers
    // hold it from then on and later launches need not touch the row; chains that never run are copied
    // once, by the row tiles behind the running ones in the launch for k = 0.  A row's arithmetic does not
    // depend on its neighbours: the chains come out bit for bit as without the ordering.
    __shared__ int rowChain[ORDERED ? kDmmaBM : 1];
    if (ORDERED) {
        if (tid < kDmmaBM) {
            const int r = c0 + tid;
            rowChain[tid] = r < chains ? f.order[r] : -1;
        }
        __syncthreads();
    }
    // chain of tile row r, or -1
    auto chainOf = [&](int r) -> int { return ORDERED ? rowChain[r] : c0 + r; };
    auto valid = [&](int c) -> bool { return ORDERED ? c >= 0 : c < chains; };      // (unordered: exactly the tests this kernel had)
    if (ORDERED && (int)blockIdx.y >= f.gemmTiles) {
        for (int r = warp * (kDmmaBM / 4); r < (warp + 1) * (kDmmaBM / 4); ++r) {
            const int c = rowChain[r];
            if (c < 0) continue;
            const size_t base = (size_t)c * n + i0;
            if (VEC16) {
                if (i0 + 2 * lane < n)
                    *reinterpret_cast<double2*>(f.qOut + base + 2 * lane) = *reinterpret_cast<const double2*>(f.qIn + base + 2 * lane);
            } else {
                if (i0 + lane < n) f.qOut[base + lane] = f.qIn[base + lane];
                if (i0 + lane + 32 < n) f.qOut[base + lane + 32] = f.qIn[base + lane + 32];
            }
        }
        return;
    }
    // the U-turn test of the PREVIOUS gradient (k - 1 >= 1), one thread per chain of this row of CTAs
    if (blockIdx.x == 0 && k >= 2 && tid < kDmmaBM && valid(chainOf(tid))) {
        const int c = chainOf(tid);
        const int st = f.leapSteps[c];
        if (st >= 1 && k - 1 <= st - 1) {
            const double* u = f.uturn + ((size_t)((k - 1) & 1) * chains + c) * f.blocks;
            double sum = 0.0;
            for (int b = 0; b < f.blocks; ++b) sum += u[b];
            if (!(sum >= 0.0)) f.okLeap[(size_t)c * f.scalarStride * 2] = 2;       // :637-638 (int index: two ints per double)
        }
    }
    double acc[4][4][2];
    dmmaMainloop<VEC16, ORDERED>(acc, dmmaSmem, f.qIn, err, c0, chains, i0, n, tid, rowChain);
    // The last gradient of a chain's trajectory is taken AT the proposed point (no drift follows,
    // :646-648): x . (Error x) over this column block is the block's share of the potential there --
    // the partial sum kDummyContractDmma (mode 1) would produce from the same accumulators, formed the
    // same way (per lane over its columns, the quad by shuffle, the two warps of a row in shared memory).
    if (f.endPartial) {
        __shared__ double part[kDmmaBM][2];
        bool mine = false;
        int rowEnd[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int c = chainOf(wm + a * 8 + g);
            const int st = valid(c) ? f.leapSteps[c] : -1;
            rowEnd[a] = (st >= 1 && k == st) ? 1 : 0;
            mine = mine || rowEnd[a];
        }
        if (__syncthreads_or(mine ? 1 : 0)) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int c = chainOf(wm + a * 8 + g);
                double sum = 0.0;
                if (rowEnd[a]) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int i = i0 + wn + b * 8 + 2 * q;
                        if (i < n) sum = fma(f.qIn[(size_t)c * n + i], acc[a][b][0], sum);
                        if (i + 1 < n) sum = fma(f.qIn[(size_t)c * n + i + 1], acc[a][b][1], sum);
                    }
                }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                if (q == 0) part[wm + a * 8 + g][warp & 1] = sum;
            }
            __syncthreads();
            if (tid < kDmmaBM && valid(chainOf(tid))) {
                const int c = chainOf(tid);
                const int st = f.leapSteps[c];
                if (st >= 1 && k == st) f.endPartial[(size_t)c * f.blocks + blockIdx.x] = part[tid][0] + part[tid][1];
            }
        }
    }
    // The gradient tile goes through shared memory (the pipeline buffers are free now) so that the
    // element-wise part reads and writes WHOLE ROWS: a warp takes a chain's 64 dimensions of this
    // column block as one 512-byte piece of q, p, p0 (16 bytes per lane), instead of the 8 x 64-byte
    // pieces of the accumulator layout.
    constexpr int kLd = kDmmaBN + 2;
    double* tile = dmmaSmem;                                   // [64][66]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            double2 v;
            v.x = acc[a][b][0];
            v.y = acc[a][b][1];
            *reinterpret_cast<double2*>(tile + (wm + a * 8 + g) * kLd + wn + b * 8 + 2 * q) = v;
        }
    __syncthreads();
    // rows in batches of kBatch: every load of a batch is issued before the first store (the row
    // loop would otherwise wait for one round trip to L2 per row: 16 in a row per warp)
    constexpr int kBatch = 4;
    int col[2];
    bool in[2];
    if (VEC16) {
        col[0] = 2 * lane;
        col[1] = 2 * lane + 1;
    } else {
        col[0] = lane;
        col[1] = lane + 32;
    }
    in[0] = i0 + col[0] < n;
    in[1] = i0 + col[1] < n;
    for (int r0 = warp * (kDmmaBM / 4); r0 < (warp + 1) * (kDmmaBM / 4); r0 += kBatch) {
        double gv[kBatch][2], qv[kBatch][2], pv[kBatch][2], p0v[kBatch][2], eps[kBatch];
        int st[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int c = chainOf(r0 + u);
            st[u] = valid(c) ? f.leapSteps[c] : -1;
            const bool live = st[u] >= 1 && k <= st[u];
            eps[u] = live ? f.epsilon[(size_t)c * f.scalarStride] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int c = chainOf(r0 + u);
            if (!valid(c)) continue;
            const bool live = st[u] >= 1 && k <= st[u];
            const bool half = (k == 0) || (k == st[u]);
            const size_t base = (size_t)c * n + i0;
            if (VEC16) {                                       // n even: both columns of a lane are in range or neither
                if (in[0]) {
                    const double2 t2 = *reinterpret_cast<const double2*>(tile + (r0 + u) * kLd + col[0]);
                    const double2 q2 = *reinterpret_cast<const double2*>(f.qIn + base + col[0]);
                    gv[u][0] = t2.x; gv[u][1] = t2.y; qv[u][0] = q2.x; qv[u][1] = q2.y;
                    if (live) {
                        const double2 p2 = *reinterpret_cast<const double2*>(f.p + base + col[0]);
                        pv[u][0] = p2.x; pv[u][1] = p2.y;
                        if (!half) {
                            const double2 z2 = *reinterpret_cast<const double2*>(f.p0 + base + col[0]);
                            p0v[u][0] = z2.x; p0v[u][1] = z2.y;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!in[h]) continue;
                    gv[u][h] = tile[(r0 + u) * kLd + col[h]];
                    qv[u][h] = f.qIn[base + col[h]];
                    if (live) {
                        pv[u][h] = f.p[base + col[h]];
                        if (!half) p0v[u][h] = f.p0[base + col[h]];
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int c = chainOf(r0 + u);
            if (!valid(c)) continue;
            const bool live = st[u] >= 1 && k <= st[u];
            const bool half = (k == 0) || (k == st[u]);
            const bool drift = live && k < st[u];
            const size_t base = (size_t)c * n + i0;
            double dot = 0.0;
            double qo[2] = {0.0, 0.0}, po[2] = {0.0, 0.0};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!in[h]) continue;
                qo[h] = qv[u][h];
                if (!live) continue;
                double kick = __dmul_rn(eps[u], gv[u][h]);
                if (half) kick = __ddiv_rn(kick, 2.0);
                po[h] = __dsub_rn(pv[u][h], kick);
                if (!half) dot += __dmul_rn(po[h], p0v[u][h]);
                if (drift) qo[h] = __dadd_rn(qv[u][h], __dmul_rn(eps[u], po[h]));
            }
            double* keep = !live ? nullptr : (k == 0 ? f.gradStart : (k == st[u] ? f.gradEnd : nullptr));
            if (VEC16) {
                if (in[0]) {
                    double2 o;
                    o.x = qo[0]; o.y = qo[1];
                    *reinterpret_cast<double2*>(f.qOut + base + col[0]) = o;
                    if (live) {
                        o.x = po[0]; o.y = po[1];
                        *reinterpret_cast<double2*>(f.p + base + col[0]) = o;
                    }
                    if (keep) {
                        o.x = gv[u][0]; o.y = gv[u][1];
                        *reinterpret_cast<double2*>(keep + base + col[0]) = o;
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!in[h]) continue;
                    f.qOut[base + col[h]] = qo[h];
                    if (live) f.p[base + col[h]] = po[h];
                    if (keep) keep[base + col[h]] = gv[u][h];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            if (lane == 0) f.uturn[((size_t)(k & 1) * chains + c) * f.blocks + blockIdx.x] = dot;
        }
    }
}

// The first leap-frog stage (k = 0) when the gradient at the starting point is already known: a step
// starts where the previous one started (rejected) or ended (accepted), and the previous step took
// the gradient at both (LeapFused::gradStart / gradEnd; kHmcAccept keeps the right one in `grad`).
// Element for element what kHmcLeapDmma's epilogue does for k = 0 -- half kick, drift, no U-turn
// term -- without the GEMM: the same values, one gradient evaluation per step less.  One warp per
// chain.
__global__ void __launch_bounds__(128)
kHmcLeapCached(LeapFused f, const double* __restrict__ grad, int chains, int n) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= chains) return;
    const int st = f.leapSteps[c];
    const bool live = st >= 1;
    const double eps = live ? f.epsilon[(size_t)c * f.scalarStride] : 0.0;
    const size_t row = (size_t)c * n;
    if ((n & 1) == 0) {
        const int m = n >> 1;
        const double2* q2 = reinterpret_cast<const double2*>(f.qIn + row);
        const double2* g2 = reinterpret_cast<const double2*>(grad + row);
        double2* p2 = reinterpret_cast<double2*>(f.p + row);
        double2* o2 = reinterpret_cast<double2*>(f.qOut + row);
        for (int i = lane; i < m; i += 64) {
            const bool two = i + 32 < m;
            double2 q[2], g[2], p[2];
            q[0] = q2[i];
            if (two) q[1] = q2[i + 32];
            if (live) {
                g[0] = g2[i]; p[0] = p2[i];
                if (two) { g[1] = g2[i + 32]; p[1] = p2[i + 32]; }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                double2 qo = q[u];
                if (live) {
                    double2 po;
                    po.x = __dsub_rn(p[u].x, __ddiv_rn(__dmul_rn(eps, g[u].x), 2.0));
                    po.y = __dsub_rn(p[u].y, __ddiv_rn(__dmul_rn(eps, g[u].y), 2.0));
                    qo.x = __dadd_rn(q[u].x, __dmul_rn(eps, po.x));
                    qo.y = __dadd_rn(q[u].y, __dmul_rn(eps, po.y));
                    p2[i + 32 * u] = po;
                }
                o2[i + 32 * u] = qo;
            }
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            double qo = f.qIn[row + i];
            if (live) {
                const double po = __dsub_rn(f.p[row + i], __ddiv_rn(__dmul_rn(eps, grad[row + i]), 2.0));
                qo = __dadd_rn(qo, __dmul_rn(eps, po));
                f.p[row + i] = po;
            }
            f.qOut[row + i] = qo;
        }
    }
}

inline void launchHmcLeapDmma(cudaStream_t stream, const double* err, const LeapFused& f, int k, int chains, int n) {
    static bool ready = false;
    if (!ready) {
        cudaFuncSetAttribute(kHmcLeapDmma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDmmaSmemBytes);
        cudaFuncSetAttribute(kHmcLeapDmma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDmmaSmemBytes);
        cudaFuncSetAttribute(kHmcLeapDmma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDmmaSmemBytes);
        cudaFuncSetAttribute(kHmcLeapDmma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDmmaSmemBytes);
        ready = true;
    }
    const int rowTiles = (chains + kDmmaBM - 1) / kDmmaBM;
    // ordered: only the row tiles of the chains that still run (and, for k = 0, the copy of the others)
    dim3 grid((n + kDmmaBN - 1) / kDmmaBN, f.order ? (k == 0 ? rowTiles : f.gemmTiles) : rowTiles);
    if (grid.y == 0) return;
    if (f.order) {
        if ((n & 1) == 0) kHmcLeapDmma<true, true><<<grid, 128, kDmmaSmemBytes, stream>>>(err, f, k, chains, n);
        else kHmcLeapDmma<false, true><<<grid, 128, kDmmaSmemBytes, stream>>>(err, f, k, chains, n);
    } else {
        if ((n & 1) == 0) kHmcLeapDmma<true, false><<<grid, 128, kDmmaSmemBytes, stream>>>(err, f, k, chains, n);
        else kHmcLeapDmma<false, false><<<grid, 128, kDmmaSmemBytes, stream>>>(err, f, k, chains, n);
    }
}

// L[c] = -1/2 sum over the column blocks of the partial sums.
__global__ void kDummyLlhFromPartials(const double* __restrict__ partial, int blocks, int m, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += partial[(size_t)c * blocks + b];
    out[c] = -0.5 * s;
}

}  // namespace smcmc
