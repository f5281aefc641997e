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
// [64][16 + pad]; rows past `rows` and columns past `n` are zero.
__device__ __forceinline__ void dmmaLoadTile(double (*dst)[kDmmaBK + kDmmaPad], const double* __restrict__ src,
                                             int row0, int rows, int k0, int n, int tid) {
    // 64 rows x 16 doubles = 1024 doubles, 128 threads: 8 each (one row, 8 consecutive k, as 8-byte copies:
    // rows of X are only 8-byte aligned when n is odd)
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

// mode 0: y[c][i] = sum_j err[i][j] x[c][j]                      (gradient of the potential)
// mode 1: partial[c][blockIdx.x] = sum_{i in this block} x[c][i] * Y[c][i]   (for the likelihood)
// leapSteps / k: as kDummyGradient (chains whose trajectory is complete are not written); may be null.
__global__ void __launch_bounds__(128)
kDummyContractDmma(const double* __restrict__ x, const double* __restrict__ err, double* __restrict__ out,
                   const int* __restrict__ leapSteps, int k, int chains, int n, int mode) {
    __shared__ __align__(16) double As[2][kDmmaBM][kDmmaBK + kDmmaPad];
    __shared__ __align__(16) double Bs[2][kDmmaBN][kDmmaBK + kDmmaPad];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;     // warp tile origin inside the CTA tile
    const int c0 = blockIdx.y * kDmmaBM, i0 = blockIdx.x * kDmmaBN;
    const int g = lane >> 2, q = lane & 3;                     // fragment coordinates
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int steps = (n + kDmmaBK - 1) / kDmmaBK;
    dmmaLoadTile(As[0], x, c0, chains, 0, n, tid);
    dmmaLoadTile(Bs[0], err, i0, n, 0, n, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int s = 0; s < steps; ++s) {
        const int cur = s & 1;
        if (s + 1 < steps) {
            dmmaLoadTile(As[cur ^ 1], x, c0, chains, (s + 1) * kDmmaBK, n, tid);
            dmmaLoadTile(Bs[cur ^ 1], err, i0, n, (s + 1) * kDmmaBK, n, tid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kDmmaBK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = As[cur][wm + a * 8 + g][kk + q];      // A(row g, col q)
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bs[cur][wn + b * 8 + g][kk + q];      // B(k q, col g) = err[i][k]
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        __syncthreads();
    }
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

// L[c] = -1/2 sum over the column blocks of the partial sums.
__global__ void kDummyLlhFromPartials(const double* __restrict__ partial, int blocks, int m, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += partial[(size_t)c * blocks + b];
    out[c] = -0.5 * s;
}

}  // namespace smcmc
