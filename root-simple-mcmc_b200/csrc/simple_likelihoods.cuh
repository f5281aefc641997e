// simple_likelihoods.cuh -- the analytic test likelihoods as device functors.
// One thread per parameter point; each sum is accumulated in the reference's
// order so the value is bit-identical to the host functor.
#pragma once
#include <cuda_runtime.h>

#include "smcmc_b200.h"

namespace smcmc {

// TSimpleMCMC.H:111-120 (documentation example): sum of -0.5*x*x.
__device__ __forceinline__ double llhUnitGauss(const double* x, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = __dadd_rn(s, __dmul_rn(__dmul_rn(-0.5, x[i]), x[i]));
    return s;
}

// TDummyLogLikelihood.H:21-31: L -= 0.5*p[i]*Error(j,i)*p[j], i outer, j inner.
__device__ __forceinline__ double llhDummy(const double* x, int n, const double* __restrict__ err) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const double hx = __dmul_rn(0.5, x[i]);
        for (int j = 0; j < n; ++j) {
            s = __dsub_rn(s, __dmul_rn(__dmul_rn(hx, err[(size_t)j * n + i]), x[j]));
        }
    }
    return s;
}

// THorrificLogLikelihood.H:26-38.
__device__ __forceinline__ double llhHorrific(const double* x, int n) {
    const double sigma = 0.01;
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        if (fabs(x[i]) > 1.0) return -1E+30;
        s = __dadd_rn(s, x[i]);
    }
    double naturalSigma = __dsqrt_rn(__ddiv_rn(__dmul_rn((double)n, 4.0), 12.0));
    s = __ddiv_rn(s, naturalSigma);
    s = __ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(-0.5, s), s), sigma), sigma);
    return s;
}

// TAsymLogLikelihood.H:20-31.
__device__ __forceinline__ double llhAsym(const double* x, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        double a = x[i];
        if (a < 0.0) a = __dmul_rn(a, 100.0);
        else a = __dmul_rn(a, -1.0);
        s = __dadd_rn(s, a);
    }
    return s;
}

// THardLogLikelihood.H:57-69 (ROSEN_B = 100).
__device__ __forceinline__ double llhHard(const double* x, int n) {
    double s = 0.0;
    for (int i = 0; i + 1 < n; ++i) {
        const double a = __dsub_rn(1.0, x[i]);
        const double b = __dsub_rn(x[i + 1], __dmul_rn(x[i], x[i]));
        s = __dsub_rn(s, __dadd_rn(__dmul_rn(a, a), __dmul_rn(__dmul_rn(100.0, b), b)));
    }
    return s;
}

// The gradient functor of THardLogLikelihood (:72-91) followed by the sign flip
// of TSimpleHMC::PotentialGradient (TSimpleHMC.H:478-487): two exact negations of
// the Rosenbrock gradient r.  One thread per (chain, dimension).
// leapSteps / k as kDummyGradient.
__global__ void kHardGradient(const double* __restrict__ x, double* __restrict__ grad,
                              const int* __restrict__ leapSteps, int k, int chains, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)chains * n) return;
    const int c = (int)(idx / n), i = (int)(idx - (size_t)c * n);
    const int steps = leapSteps[c];
    if (steps < 1 || k > steps) return;
    const double* p = x + (size_t)c * n;
    double g;
    if (i == 0) {
        g = __dsub_rn(-__dmul_rn(2.0, __dsub_rn(1.0, p[0])),
                      __dmul_rn(__dmul_rn(400.0, p[0]), __dsub_rn(p[1], __dmul_rn(p[0], p[0]))));
    } else if (i < n - 1) {
        g = __dmul_rn(200.0, __dsub_rn(p[i], __dmul_rn(p[i - 1], p[i - 1])));
        g = __dadd_rn(g, -__dmul_rn(2.0, __dsub_rn(1.0, p[i])));
        g = __dadd_rn(g, __dmul_rn(__dmul_rn(-400.0, p[i]), __dsub_rn(p[i + 1], __dmul_rn(p[i], p[i]))));
    } else {
        g = __dmul_rn(200.0, __dsub_rn(p[i], __dmul_rn(p[i - 1], p[i - 1])));
    }
    grad[idx] = -(-g);
}

// TDummyLogLikelihood for many points: one THREAD per point, the block's points
// transposed into shared memory (xs[j][point], lane = point: conflict-free).  The
// error matrix is read through its transpose errT[i][j] = Error(j,i), one row per
// outer index i, copied into shared memory one row ahead with cp.async so that
// the inner loop only touches shared memory.  The single running sum of the
// reference (n^2 terms, i outer, j inner) stays a single sequential sum per
// point; the two products of each term are independent work beside that chain.
// Dynamic shared memory: n * blockDim.x + 2 * (n rounded up to 2) doubles.
__device__ __forceinline__ void cpAsyncRow(double* dst, const double* src, int n, int tid, int nthreads) {
    // 16-byte pieces where the row is 16-byte aligned, 8-byte pieces otherwise
    const bool aligned = ((reinterpret_cast<size_t>(src) | reinterpret_cast<size_t>(dst)) & 15) == 0;
    if (aligned) {
        for (int k = tid * 2; k + 1 < n; k += nthreads * 2)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + k)),
                         "l"(src + k) : "memory");
        if ((n & 1) && tid == 0)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst + n - 1)),
                         "l"(src + n - 1) : "memory");
    } else {
        for (int k = tid; k < n; k += nthreads)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst + k)),
                         "l"(src + k) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void kDummyLikelihood(const double* __restrict__ x, int m, int n, const double* __restrict__ errT,
                                 double* __restrict__ out) {
    extern __shared__ __align__(16) double xs[];
    const int width = blockDim.x;
    const int npad = (n + 1) & ~1;
    double* rows = xs + (size_t)n * width;          // two row buffers of npad doubles
    const int c0 = blockIdx.x * width;
    cpAsyncRow(rows, errT, n, threadIdx.x, width);
    for (int idx = threadIdx.x; idx < n * width; idx += width) {
        const int j = idx / width, c = idx - j * width;
        xs[idx] = (c0 + c < m) ? x[(size_t)(c0 + c) * n + j] : 0.0;
    }
    const int c = threadIdx.x;
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const double* row = rows + (size_t)(i & 1) * npad;
        if (i + 1 < n) cpAsyncRow(rows + (size_t)((i + 1) & 1) * npad, errT + (size_t)(i + 1) * n, n, threadIdx.x, width);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");     // row i has landed (row i+1 may be in flight)
        __syncthreads();
        const double hx = __dmul_rn(0.5, xs[i * width + c]);
#pragma unroll 4
        for (int j = 0; j < n; ++j) s = __dsub_rn(s, __dmul_rn(__dmul_rn(hx, row[j]), xs[j * width + c]));
        __syncthreads();                                         // before row buffer (i & 1) is refilled
    }
    if (c0 + c < m) out[c0 + c] = s;
}

__global__ void kSimpleLikelihood(int kind, const double* __restrict__ x, int m, int n,
                                  const double* __restrict__ err, double* __restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const double* p = x + (size_t)c * n;
    double v;
    switch (kind) {
    case SMCMC_LLH_UNIT_GAUSS: v = llhUnitGauss(p, n); break;
    case SMCMC_LLH_DUMMY: v = llhDummy(p, n, err); break;
    case SMCMC_LLH_HORRIFIC: v = llhHorrific(p, n); break;
    case SMCMC_LLH_HARD: v = llhHard(p, n); break;
    default: v = llhAsym(p, n); break;
    }
    out[c] = v;
}

}  // namespace smcmc
