// constrained_functor.cuh -- example4's likelihood (reference
// example4/TConstrainedLikelihood.H:26-46, priors of its Init() :55-110) written
// as a USER DEVICE FUNCTOR for include/smcmc_device_functor.cuh: a trivially
// copyable class with a __host__ __device__ call operator on a plain array.
// This is the user's side of the plugin contract -- nothing in libsmcmc_b200
// knows this likelihood.  The arithmetic keeps the reference's operation order
// (compile with -fmad=false for bit parity with the host code).
#ifndef TESTS_CONSTRAINED_FUNCTOR_CUH
#define TESTS_CONSTRAINED_FUNCTOR_CUH

#include "smcmc_device_functor.cuh"

class TConstrainedLikelihood {
public:
    static const int kMaxDim = 32;
    double ExpectedValues[kMaxDim];
    double PriorConstraints[kMaxDim];
    double SummedValues;
    double SummedConstraint;
    int Dimensions;

    TConstrainedLikelihood() : SummedValues(0.0), SummedConstraint(1.0), Dimensions(0) {}

    std::size_t GetDim() const { return Dimensions; }

    // log(likelihood): the sum is constrained, and so is every value
    __host__ __device__ double operator()(const double* point, int n) const {
        double logLikelihood = 0.0;
        double sum = 0.0;
        for (int i = 0; i < n; ++i) sum += point[i];
        sum = (sum - SummedValues) / SummedConstraint;
        logLikelihood -= 0.5 * sum * sum;
        for (int i = 0; i < n; ++i) {
            double v = point[i] - ExpectedValues[i];
            v /= PriorConstraints[i];
            logLikelihood -= 0.5 * v * v;
        }
        return logLikelihood;
    }
    double operator()(const sMCMC::Vector& point) const { return (*this)(point.data(), (int)point.size()); }

    // grad(log(likelihood)); the reference's functor declines (:49-51), this one is
    // written out so that TSimpleHMC<L, L> has a user gradient to call
    __host__ __device__ bool Gradient(const double* point, int n, double* g) const {
        double sum = 0.0;
        for (int i = 0; i < n; ++i) sum += point[i];
        const double common = (sum - SummedValues) / SummedConstraint / SummedConstraint;
        for (int i = 0; i < n; ++i) {
            const double v = (point[i] - ExpectedValues[i]) / PriorConstraints[i] / PriorConstraints[i];
            g[i] = -common - v;
        }
        return true;
    }

    // the priors of the reference's Init()
    void Init() {
        SummedValues = 1902.0;
        SummedConstraint = 16.0;
        Dimensions = 25;
        for (int i = 0; i < 24; ++i) {
            ExpectedValues[i] = 76.0;
            PriorConstraints[i] = 76.0 * 0.08;
        }
        ExpectedValues[24] = 80.0;
        PriorConstraints[24] = 2.0;
    }
};

#endif
