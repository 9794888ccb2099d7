// Shared helpers for the TASU bridge kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/tasu_bridge.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtasu_bridge is written for sm_100a (B200) only"
#endif

namespace tasu {

void set_error(const char* fmt, ...);

#define TASU_CHECK_ARG(cond, msg)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            tasu::set_error("%s: invalid argument: %s", __func__, msg); \
            return TASU_ERR_INVALID_ARG;                            \
        }                                                           \
    } while (0)

#define TASU_CHECK_CUDA(expr)                                                          \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            tasu::set_error("%s: CUDA error %s at %s:%d", __func__, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return TASU_ERR_CUDA;                                                      \
        }                                                                              \
    } while (0)

#define TASU_CHECK_LAUNCH() TASU_CHECK_CUDA(cudaGetLastError())

int sm_count();
int option(int id);          // tasu_set_option / tasu_get_option values (core.cu)

// grid of a persistent (grid-stride) kernel: SMs x CTAs that are actually resident for this kernel, capped by the
// number of work items — a larger grid only adds a ragged second wave
template <typename K>
inline unsigned persistent_grid(K kernel, int block, int64_t items) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm <= 0) per_sm = 4;
    int64_t g = (int64_t)sm_count() * per_sm;
    if (g > items) g = items;
    return (unsigned)(g < 1 ? 1 : g);
}

constexpr int kWarp = 32;

// ---- order-preserving float <-> uint32 (so atomicMax on uint gives float max; 0 < every float)
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

// ---- dtype helpers
template <typename T> struct Vec16;                       // 16-byte vector of T
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// streaming 128-bit load that does not pollute L1 (data is read exactly once)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// unpack a 16-byte register quad into floats
__device__ __forceinline__ void unpack16(const uint4& q, float (&f)[4], float) {
    f[0] = __uint_as_float(q.x); f[1] = __uint_as_float(q.y);
    f[2] = __uint_as_float(q.z); f[3] = __uint_as_float(q.w);
}
__device__ __forceinline__ void unpack16(const uint4& q, float (&f)[8], __nv_bfloat16) {
    // bf16 -> fp32 is a 16-bit shift
    f[0] = __uint_as_float(q.x << 16); f[1] = __uint_as_float(q.x & 0xffff0000u);
    f[2] = __uint_as_float(q.y << 16); f[3] = __uint_as_float(q.y & 0xffff0000u);
    f[4] = __uint_as_float(q.z << 16); f[5] = __uint_as_float(q.z & 0xffff0000u);
    f[6] = __uint_as_float(q.w << 16); f[7] = __uint_as_float(q.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// block-wide exclusive scan of one int per thread (blockDim.x multiple of 32, <= 1024).
// `scratch` needs 33 ints of shared memory. Returns exclusive prefix; *total = block sum.
__device__ __forceinline__ int block_excl_scan_i(int v, int* scratch, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int incl = warp_incl_scan_i(v, lane);
    __syncthreads();                       // protect scratch from a previous use
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarp ? scratch[lane] : 0;
        int wi = warp_incl_scan_i(w, lane);
        scratch[lane] = wi - w;            // exclusive warp offsets
        if (lane == 31) scratch[32] = wi;
    }
    __syncthreads();
    *total = scratch[32];
    return incl - v + scratch[warp];
}

// block-wide sum / max of ints and floats (result broadcast to all threads)
__device__ __forceinline__ int block_sum_i(int v, int* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    v = warp_sum_i(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int r = 0;
    for (int i = 0; i < nwarp; ++i) r += scratch[i];
    return r;
}
__device__ __forceinline__ int block_max_i(int v, int* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    v = warp_max_i(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int r = scratch[0];
    for (int i = 1; i < nwarp; ++i) r = max(r, scratch[i]);
    return r;
}
__device__ __forceinline__ float block_sum_f(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < nwarp; ++i) r += scratch[i];
    return r;
}

}  // namespace tasu
