// common.cuh -- error handling and device-buffer plumbing for libsmcmc_b200.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "smcmc_b200.h"

namespace smcmc {

struct Error : public std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

inline void cudaCheck(cudaError_t e, const char* expr, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e), file, line, expr);
        throw Error(SMCMC_ERR_CUDA, buf);
    }
}
#define CUDA_CHECK(expr) ::smcmc::cudaCheck((expr), #expr, __FILE__, __LINE__)

// Device memory comes from the device's default stream-ordered pool
// (cudaMallocAsync) with a release threshold of 4 GiB: an engine that is created,
// loaded and destroyed again (one per sample, as the end-to-end measurement of
// bench.py does) gets its buffers back from the pool instead of paying cudaMalloc /
// cudaFree: a 1M-event upload went from 16.6 ms to 1.1 ms on B200 (bench.py e2e).  The
// ordering is kept as strict as cudaFree's: a release synchronises the device
// first, an allocation is complete when reserve() returns.  SMCMC_NO_POOL=1 uses
// cudaMalloc / cudaFree.
inline bool devicePoolEnabled() {
    static const bool on = [] {
        if (std::getenv("SMCMC_NO_POOL")) return false;
        int dev = 0, supported = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev);
        return supported != 0;
    }();
    return on;
}
inline void devicePoolConfigure(int dev) {
    static bool done[64] = {false};
    if (dev < 0 || dev >= 64 || done[dev] || !devicePoolEnabled()) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t threshold = 4ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    done[dev] = true;
}
inline cudaError_t deviceAlloc(void** p, size_t bytes) {
    if (!devicePoolEnabled()) return cudaMalloc(p, bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    devicePoolConfigure(dev);
    cudaError_t e = cudaMallocAsync(p, bytes, cudaStreamLegacy);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(cudaStreamLegacy);
}
inline void deviceFree(void* p) {
    if (!devicePoolEnabled()) {
        cudaFree(p);
        return;
    }
    cudaDeviceSynchronize();
    cudaFreeAsync(p, cudaStreamLegacy);
}

// A typed device allocation, grow-only.
template <class T>
class DeviceBuffer {
public:
    DeviceBuffer() : ptr_(nullptr), count_(0) {}
    ~DeviceBuffer() { release(); }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    void release() {
        if (ptr_) deviceFree(ptr_);
        ptr_ = nullptr;
        count_ = 0;
    }
    // Ensure room for n elements (contents undefined after growth).
    void reserve(size_t n) {
        if (n <= count_) return;
        release();
        CUDA_CHECK(deviceAlloc((void**)&ptr_, n * sizeof(T)));
        count_ = n;
    }
    void swap(DeviceBuffer& o) {
        T* p = ptr_; ptr_ = o.ptr_; o.ptr_ = p;
        size_t c = count_; count_ = o.count_; o.count_ = c;
    }
    T* get() const { return ptr_; }
    size_t count() const { return count_; }
    size_t bytes() const { return count_ * sizeof(T); }
private:
    T* ptr_;
    size_t count_;
};

inline int ceilDiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace smcmc
