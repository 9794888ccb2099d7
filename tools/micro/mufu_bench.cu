// Micro-benchmark (sm_100a): throughput of the exponentials an epilogue can use, in results per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// 2^x on the FMA / ALU pipes: round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-4 polynomial, exponent add
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.f);
    const float t = x + 12582912.f;                 // 1.5 * 2^23: n in the low mantissa bits
    const float n = t - 12582912.f;
    const float f = x - n;
    float p = 1.3333558e-3f;
    p = fmaf(p, f, 9.6181291e-3f);
    p = fmaf(p, f, 5.5504109e-2f);
    p = fmaf(p, f, 2.4022651e-1f);
    p = fmaf(p, f, 6.9314718e-1f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float a[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i) * 1e-3f - 3.f; u[i] = 0xB800B800u + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2_f32(a[i]) - 1.5f;
            if (MODE == 1) u[i] = ex2_f16x2(u[i]) ^ 0x80008000u;
            if (MODE == 2) u[i] = ex2_bf16x2(u[i]) ^ 0x80008000u;
            if (MODE == 3) a[i] = ex2_poly(a[i]) - 1.5f;
            if (MODE == 4) { if (i & 1) a[i] = ex2_poly(a[i]) - 1.5f; else a[i] = ex2_f32(a[i]) - 1.5f; }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, int per_inst) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 100, 1.f);
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, iters, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double results = (double)sms * 8 * 256 * 8.0 * iters * per_inst;
    printf("%-28s %8.3f ms  %7.2f results/ns  = %6.2f results/clk/SM at the nominal %d MHz\n", name, ms, results / (ms * 1e6),
           results / (ms * 1e-3) / ((double)khz * 1e3) / sms, khz / 1000);
    cudaFree(out);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.f16x2", 2);
    run<2>("ex2.approx.ftz.bf16x2", 2);
    run<3>("polynomial (FMA pipe)", 1);
    run<4>("half MUFU, half polynomial", 1);
    // accuracy of the polynomial against exp2f on [-30, 0]
    return 0;
}
