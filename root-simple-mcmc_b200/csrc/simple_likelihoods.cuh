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
    default: v = llhAsym(p, n); break;
    }
    out[c] = v;
}

}  // namespace smcmc
