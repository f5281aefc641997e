// common.cuh -- error handling and device-buffer plumbing for libsmcmc_b200.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
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

// A typed device allocation (cudaMalloc / cudaFree), grow-only.
template <class T>
class DeviceBuffer {
public:
    DeviceBuffer() : ptr_(nullptr), count_(0) {}
    ~DeviceBuffer() { release(); }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    void release() {
        if (ptr_) cudaFree(ptr_);
        ptr_ = nullptr;
        count_ = 0;
    }
    // Ensure room for n elements (contents undefined after growth).
    void reserve(size_t n) {
        if (n <= count_) return;
        release();
        CUDA_CHECK(cudaMalloc((void**)&ptr_, n * sizeof(T)));
        count_ = n;
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
