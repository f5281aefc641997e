// smcmc_device_functor.cuh -- user-written likelihood (and gradient) functors
// for the B200 engine: the reference's plugin contract, on the device.
//
// In root-simple-mcmc the whole plugin API is "hand TSimpleMCMC your own
// functor" (reference TSimpleMCMC.H:48-106, :185-187, :544; used that way in
// example/FakeMCMC.C:28-30 and example4/Constrained.C:17-25):
//
//     struct MyLikelihood { double operator()(const sMCMC::Vector& point); };
//     sMCMC::TSimpleMCMC<MyLikelihood> mcmc(tree);
//
// Here the Step hot path runs on the GPU, so the functor's arithmetic has to be
// device code.  A user translation unit compiled by nvcc (-arch=sm_100a)
// includes THIS header (before or instead of TSimpleMCMC.H / TSimpleHMC.H),
// gives the functor a device call operator on a plain array,
//
//     struct MyLikelihood {
//         __host__ __device__ double operator()(const double* x, int n) const;
//         double operator()(const sMCMC::Vector& p) const { return (*this)(p.data(), (int)p.size()); }
//         // optional, for TSimpleHMC<MyLikelihood, MyLikelihood>: the gradient of log L
//         // (reference contract TSimpleHMC.H:38-60: true iff it was computed)
//         __host__ __device__ bool Gradient(const double* x, int n, double* g) const;
//     };
//
// and uses sMCMC::TSimpleMCMC<MyLikelihood> exactly as with the reference header.
// The functor object must be trivially copyable (plain data members, fixed-size
// arrays): it is copied by value to the device, as the reference holds it by
// value (TSimpleMCMC.H:544).  Change its members through GetLogLikelihood() before
// Start(), or call SyncLikelihood() afterwards.
//
// What this header instantiates in the user's translation unit, from the
// functor's type: the evaluation kernels below and two host launch functions,
// registered with the engine as its smcmc_user_ops table entry
// (smcmc_user_set_ops, SMCMC_LLH_USER).  No device function pointer crosses a
// module boundary: libsmcmc_b200.so calls the HOST launch function, which
// queues the kernel compiled here on the engine's stream.
#ifndef SMCMC_DEVICE_FUNCTOR_CUH_SEEN
#define SMCMC_DEVICE_FUNCTOR_CUH_SEEN

#ifndef __CUDACC__
#error "smcmc_device_functor.cuh must be compiled by nvcc (user device functors are CUDA code)"
#endif

#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "TSimpleMCMC.H"

namespace smcmc_user {

constexpr int kThreads = 128;            // points per CTA (one thread per point)
constexpr int kMaxStageBytes = 96 * 1024;

// One THREAD per point.  The kThreads rows of a CTA are one contiguous block of
// the chain-major point array: the CTA copies it to shared memory with coalesced
// loads (row stride ld = n rounded up to odd: lane-per-row reads are then
// conflict-free) and hands every thread a pointer to its own row.  Rows too long
// to stage are read in place.
template <class F>
__global__ void __launch_bounds__(kThreads)
kLikelihood(const F* __restrict__ f, const double* __restrict__ x, int m, int n, int ld, double* __restrict__ out) {
    extern __shared__ double tile[];
    const int base = blockIdx.x * kThreads;
    const int rows = min(kThreads, m - base);
    const int p = base + threadIdx.x;
    if (ld > 0) {
        const double* src = x + (size_t)base * n;
        for (int k = threadIdx.x; k < rows * n; k += kThreads) {
            const int r = k / n;
            tile[r * ld + (k - r * n)] = src[k];
        }
        __syncthreads();
        if (p < m) out[p] = (*f)(tile + threadIdx.x * ld, n);
    } else if (p < m) {
        out[p] = (*f)(x + (size_t)p * n, n);
    }
}

// Gradient of log L at every point that still integrates its trajectory, negated
// (TSimpleHMC.H:478-487: the potential is -log L).  The point and its gradient are
// staged in shared memory when they fit: [rows][ld] points, then [rows][ld] gradients.
template <class F>
__global__ void __launch_bounds__(kThreads)
kGradient(const F* __restrict__ f, const double* __restrict__ x, int m, int n, int ld, double* __restrict__ grad,
          const int32_t* __restrict__ steps, int k) {
    extern __shared__ double tile[];
    const int base = blockIdx.x * kThreads;
    const int rows = min(kThreads, m - base);
    const int p = base + threadIdx.x;
    bool live = p < m;
    if (live && steps) {
        const int s = steps[p];
        live = s >= 1 && k <= s;
    }
    if (ld > 0) {
        double* gt = tile + kThreads * ld;
        const double* src = x + (size_t)base * n;
        for (int q = threadIdx.x; q < rows * n; q += kThreads) {
            const int r = q / n;
            tile[r * ld + (q - r * n)] = src[q];
        }
        __syncthreads();
        bool ok = false;
        if (live) ok = f->Gradient(tile + threadIdx.x * ld, n, gt + threadIdx.x * ld);
        if (live && !ok)
            for (int i = 0; i < n; ++i) gt[threadIdx.x * ld + i] = nan("");   // a functor that declines at run time
        __syncthreads();
        for (int q = threadIdx.x; q < rows * n; q += kThreads) {
            const int r = q / n;
            bool rl = true;
            if (steps) {
                const int s = steps[base + r];
                rl = s >= 1 && k <= s;
            }
            if (rl) grad[(size_t)base * n + q] = -gt[r * ld + (q - r * n)];
        }
    } else if (live) {
        double* g = grad + (size_t)p * n;
        const bool ok = f->Gradient(x + (size_t)p * n, n, g);
        for (int i = 0; i < n; ++i) g[i] = ok ? -g[i] : nan("");
    }
}

template <class F, class = void>
struct HasGradient : std::false_type {};
template <class F>
struct HasGradient<F, decltype((void)std::declval<const F&>().Gradient((const double*)0, 0, (double*)0))> : std::true_type {};

// The device copy of one functor object and its launch table.
template <class F>
class Binding {
    static_assert(std::is_trivially_copyable<F>::value,
                  "a device likelihood functor is copied to the GPU by value: it must be trivially copyable "
                  "(plain data members and fixed-size arrays, no std::vector)");

public:
    Binding() : fDevice(nullptr) {}
    ~Binding() {
        if (fDevice) cudaFree(fDevice);
    }
    Binding(const Binding&) = delete;
    Binding& operator=(const Binding&) = delete;

    /// Copy the functor to the GPU (again).
    void Upload(const F& host) {
        if (!fDevice) Check(cudaMalloc((void**)&fDevice, sizeof(F)));
        Check(cudaMemcpy(fDevice, &host, sizeof(F), cudaMemcpyHostToDevice));
    }
    /// Upload and register with an engine created with SMCMC_LLH_USER.
    void Bind(smcmc_engine* e, const F& host) {
        Upload(host);
        smcmc_user_ops ops;
        ops.struct_size = sizeof ops;
        ops.reserved_ = 0;
        ops.ctx = this;
        ops.likelihood = &Binding::LaunchLikelihood;
        ops.gradient = GradientEntry(HasGradient<F>());
        const int rc = smcmc_user_set_ops(e, &ops);
        if (rc != SMCMC_OK) throw std::logic_error(smcmc_last_error(e));
    }
    static bool kHasGradient() { return HasGradient<F>::value; }

private:
    typedef int (*GradFn)(void*, const double*, int, int, double*, const int32_t*, int, void*);
    static GradFn GradientEntry(std::true_type) { return &Binding::LaunchGradient; }
    static GradFn GradientEntry(std::false_type) { return nullptr; }

    static void Check(cudaError_t rc) {
        if (rc != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(rc));
    }
    // shared-memory row stride (0: rows are read in place)
    static int Stride(int n, int copies) {
        const int ld = n | 1;
        return ((size_t)copies * kThreads * ld * sizeof(double) <= (size_t)kMaxStageBytes) ? ld : 0;
    }
    static int LaunchLikelihood(void* ctx, const double* x, int m, int n, double* out, void* stream) {
        Binding* b = static_cast<Binding*>(ctx);
        const int ld = Stride(n, 1);
        const size_t smem = (size_t)kThreads * ld * sizeof(double);
        if (smem > 48 * 1024) {
            cudaError_t rc = cudaFuncSetAttribute(kLikelihood<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (rc != cudaSuccess) return (int)rc;
        }
        kLikelihood<F><<<(m + kThreads - 1) / kThreads, kThreads, smem, (cudaStream_t)stream>>>(b->fDevice, x, m, n, ld, out);
        return (int)cudaGetLastError();
    }
    static int LaunchGradient(void* ctx, const double* x, int m, int n, double* grad, const int32_t* steps, int k,
                              void* stream) {
        Binding* b = static_cast<Binding*>(ctx);
        const int ld = Stride(n, 2);
        const size_t smem = (size_t)2 * kThreads * ld * sizeof(double);
        if (smem > 48 * 1024) {
            cudaError_t rc = cudaFuncSetAttribute(kGradient<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (rc != cudaSuccess) return (int)rc;
        }
        kGradient<F><<<(m + kThreads - 1) / kThreads, kThreads, smem, (cudaStream_t)stream>>>(b->fDevice, x, m, n, ld, grad,
                                                                                            steps, k);
        return (int)cudaGetLastError();
    }

    F* fDevice;
};

}  // namespace smcmc_user

// The sampler side: a likelihood class WITHOUT a built-in id (no kDeviceLikelihood
// member) is a user device functor.
namespace sMCMC {
namespace detail {
template <class L, class Enable>
struct DeviceBinding {
    static int Kind() { return SMCMC_LLH_USER; }
    static bool HasGradient() { return smcmc_user::HasGradient<L>::value; }
    void Bind(smcmc_engine* e, L& like) { fBinding.Bind(e, like); }
    void Sync(smcmc_engine*, L& like) { fBinding.Upload(like); }
    smcmc_user::Binding<L> fBinding;
};
}  // namespace detail
}  // namespace sMCMC

#endif
