// Micro-benchmarks behind the pair-kernel design:
// warp-instruction issue rates of FFMA, FFMA2 (packed f32x2), MUFU.EX2,
// conflict-free shared-memory read-modify-write and shared atomics on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue issue.cu && ./issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(uint64_t v) { float a, b; asm("mov.b64 {%0,%1},%2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

__global__ void kFfma(float* out, float a, float b) {
    float v[8];
    for (int k = 0; k < 8; ++k) v[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], a, b);
    float s = 0; for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 1.2345f) out[0] = s;
}
__global__ void kFfma2(float* out, float a, float b) {
    uint64_t v[8], A = pk(a, a), B = pk(b, b);
    for (int k = 0; k < 8; ++k) v[k] = pk(threadIdx.x + k, k);
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32x2 %0,%0,%1,%2;" : "+l"(v[k]) : "l"(A), "l"(B));
    float s = 0; for (int k = 0; k < 8; ++k) s += lo(v[k]);
    if (s == 1.2345f) out[0] = s;
}
// 4 FFMA2 + 4 integer IMAD per iteration: do the packed ops leave issue slots for the ALU pipe?
__global__ void kFfma2Imad(float* out, float a, float b, int m) {
    uint64_t v[4], A = pk(a, a), B = pk(b, b);
    int w[4];
    for (int k = 0; k < 4; ++k) { v[k] = pk(threadIdx.x + k, k); w[k] = threadIdx.x + k; }
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            asm volatile("fma.rn.f32x2 %0,%0,%1,%2;" : "+l"(v[k]) : "l"(A), "l"(B));
            asm volatile("lop3.b32 %0,%0,%1,%2,0x96;" : "+r"(w[k]) : "r"(m), "r"(i));
        }
    float s = 0; for (int k = 0; k < 4; ++k) s += lo(v[k]) + w[k];
    if (s == 1.2345f) out[0] = s;
}
__global__ void kFfmaLop(float* out, float a, float b, int m) {
    float v[4]; int w[4];
    for (int k = 0; k < 4; ++k) { v[k] = threadIdx.x + k; w[k] = threadIdx.x + k; }
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            asm volatile("fma.rn.f32 %0,%0,%1,%2;" : "+f"(v[k]) : "f"(a), "f"(b));
            asm volatile("lop3.b32 %0,%0,%1,%2,0x96;" : "+r"(w[k]) : "r"(m), "r"(i));
        }
    float s = 0; for (int k = 0; k < 4; ++k) s += v[k] + w[k];
    if (s == 1.2345f) out[0] = s;
}
__global__ void kMufu(float* out, float a) {
    float v[8];
    for (int k = 0; k < 8; ++k) v[k] = a * (threadIdx.x + k);
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0,%0;" : "+f"(v[k]));
    float s = 0; for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 1.2345f) out[0] = s;
}
// MUFU + 3 FFMA per MUFU: does the XU pipe overlap with FMA issue?
__global__ void kMufuFfma(float* out, float a, float b) {
    float v[4], w[4];
    for (int k = 0; k < 4; ++k) { v[k] = a * (threadIdx.x + k); w[k] = k; }
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            asm volatile("ex2.approx.ftz.f32 %0,%0;" : "+f"(v[k]));
            w[k] = fmaf(w[k], a, b); w[k] = fmaf(w[k], a, b); w[k] = fmaf(w[k], a, b);
        }
    float s = 0; for (int k = 0; k < 4; ++k) s += v[k] + w[k];
    if (s == 1.2345f) out[0] = s;
}
// private 16-bit counters: LDS.U16 / IADD / STS.U16 at a data-dependent row
__global__ void kRmw16(float* out, int m) {
    __shared__ uint16_t c[32 * 256];
    for (int i = threadIdx.x; i < 32 * 256; i += blockDim.x) c[i] = 0;
    __syncthreads();
    unsigned r = threadIdx.x * 7u + m;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            r = r * 1664525u + 1013904223u;
            unsigned row = r >> 27;
            c[row * 256 + threadIdx.x] += 1;
        }
    }
    __syncthreads();
    if (c[threadIdx.x] == 65535 && m == 12345) out[0] = 1;
}
// the same with a 32-bit shared atomic (RED-like, result unused)
__global__ void kAtom32(float* out, int m) {
    __shared__ unsigned c[32 * 256];
    for (int i = threadIdx.x; i < 32 * 256; i += blockDim.x) c[i] = 0;
    __syncthreads();
    unsigned r = threadIdx.x * 7u + m;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            r = r * 1664525u + 1013904223u;
            unsigned row = r >> 27;
            atomicAdd(&c[row * 256 + threadIdx.x], 1u);
        }
    }
    __syncthreads();
    if (c[threadIdx.x] == 65535 && m == 12345) out[0] = 1;
}
// LCG only (to subtract from the two above)
__global__ void kLcg(float* out, int m) {
    unsigned r = threadIdx.x * 7u + m, acc = 0;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { r = r * 1664525u + 1013904223u; acc += (r >> 26) * 256; }
    }
    if (acc == 12345u && m == 12345) out[0] = 1;
}

template <class F>
static double timeIt(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch();
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 4);
    const int sms = p.multiProcessorCount;
    const int blocks = sms * 4, threads = 256;           // 32 warps per SM
    const double warps = (double)blocks * threads / 32;
    auto rate = [&](double ms, double instPerThread) {   // warp-instructions per SM per clock (nominal boost clock)
        return warps * instPerThread / (ms * 1e-3) / sms / (clk * 1e3);
    };
    printf("%s, %d SMs, %d kHz\n", p.name, sms, clk);
    double t;
    t = timeIt([&] { kFfma<<<blocks, threads>>>(out, 0.999f, 1e-3f); });
    printf("FFMA            %8.3f ms  %.2f warp-inst/clk/SM\n", t, rate(t, 8.0 * ITERS));
    t = timeIt([&] { kFfma2<<<blocks, threads>>>(out, 0.999f, 1e-3f); });
    printf("FFMA2           %8.3f ms  %.2f warp-inst/clk/SM (x2 FMAs each)\n", t, rate(t, 8.0 * ITERS));
    t = timeIt([&] { kFfmaLop<<<blocks, threads>>>(out, 0.999f, 1e-3f, 5); });
    printf("FFMA+LOP3       %8.3f ms  %.2f warp-inst/clk/SM\n", t, rate(t, 8.0 * ITERS));
    t = timeIt([&] { kFfma2Imad<<<blocks, threads>>>(out, 0.999f, 1e-3f, 5); });
    printf("FFMA2+LOP3      %8.3f ms  %.2f warp-inst/clk/SM\n", t, rate(t, 8.0 * ITERS));
    t = timeIt([&] { kMufu<<<blocks, threads>>>(out, 1e-3f); });
    printf("MUFU.EX2        %8.3f ms  %.2f warp-inst/clk/SM\n", t, rate(t, 8.0 * ITERS));
    t = timeIt([&] { kMufuFfma<<<blocks, threads>>>(out, 0.999f, 1e-3f); });
    printf("MUFU+3FFMA      %8.3f ms  %.2f warp-inst/clk/SM\n", t, rate(t, 16.0 * ITERS));
    double tl = timeIt([&] { kLcg<<<blocks, threads>>>(out, 5); });
    printf("LCG only        %8.3f ms\n", tl);
    t = timeIt([&] { kRmw16<<<blocks, threads>>>(out, 5); });
    printf("LDS/IADD/STS.16 %8.3f ms  (%.3f ms over LCG) %.2f updates/clk/SM\n", t, t - tl, rate(t - tl, 8.0 * ITERS) * 32);
    t = timeIt([&] { kAtom32<<<blocks, threads>>>(out, 5); });
    printf("ATOMS.ADD.32    %8.3f ms  (%.3f ms over LCG) %.2f updates/clk/SM\n", t, t - tl, rate(t - tl, 8.0 * ITERS) * 32);
    return 0;
}
