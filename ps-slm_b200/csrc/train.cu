// Training-side helpers of the linear-silu projector (backward of Multitask/model/projector.py:149-151,
// which the reference leaves to autograd; trained in Multitask/utils/deepspeed_utils.py:235-236).
// The three large contractions (dW2, dh, G = (rstd·dz)^T·x) run on the tcgen05 GEMM; these kernels
// produce its K-major operands (transposes), the SiLU forward/backward and the LayerNorm-fold
// algebra that turns G into dW1 / dgamma / dbeta without the reference's third big GEMM
// (dLN = dh1·W1): with x̂ = rstd·(x − μ),
//   dW1[j,v]  = γ[v]·(G[j,v] − g0[j]) + β[v]·db1[j],   G = Σ_n rstd_n dz[n,j] x[n,v],  g0[j] = Σ_n rstd_n μ_n dz[n,j]
//   dγ[v]     = Σ_j W1[j,v]·(G[j,v] − g0[j])
//   dβ[v]     = Σ_j W1[j,v]·db1[j],      db1[j] = Σ_n dz[n,j]
#include "common.cuh"

namespace tasu {

// dst[c, r] = bf16(scale[r] * src[r, c]); 32x32 tiles through shared memory, coalesced both ways
template <typename Ti>
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const Ti* __restrict__ src, int64_t R, int64_t C, int64_t sstride,
                      const float* __restrict__ scale, __nv_bfloat16* __restrict__ dst, int64_t dstride) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
        float v = 0.f;
        if (r < R && c < C) {
            v = to_f32(src[r * sstride + c]);
            if (scale) v *= scale[r];
        }
        tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
        if (c < C && r < R) dst[c * dstride + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
}

__global__ void __launch_bounds__(256)
silu_fwd_kernel(const float* __restrict__ z, int64_t n, __nv_bfloat16* __restrict__ h) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(z + i);
        const float a = v.x / (1.f + __expf(-v.x)), b = v.y / (1.f + __expf(-v.y));
        const float c = v.z / (1.f + __expf(-v.z)), d = v.w / (1.f + __expf(-v.w));
        *reinterpret_cast<uint2*>(h + i) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
    } else {
        for (int64_t k = i; k < n; ++k) { const float x = z[k]; h[k] = __float2bfloat16_rn(x / (1.f + __expf(-x))); }
    }
}

// dz = dh * silu'(z); dzsT[j, n] = bf16(rstd[n] * dz[n, j]); db1[j] += Σ_n dz; g0[j] += Σ_n rstd μ dz
__global__ void __launch_bounds__(256)
silu_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ z, int64_t N, int Hb,
                const float* __restrict__ rstd, const float* __restrict__ mean,
                __nv_bfloat16* __restrict__ dzsT, int64_t tstride, float* __restrict__ db1, float* __restrict__ g0) {
    __shared__ float tile[32][33];
    __shared__ float s_db[8][32], s_g0[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t n0 = (int64_t)blockIdx.y * 32;
    const int j0 = blockIdx.x * 32;
    float a_db = 0.f, a_g0 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t n = n0 + ty + 8 * i;
        const int j = j0 + tx;
        float v = 0.f;
        if (n < N && j < Hb) {
            const float zz = z[n * Hb + j], s = 1.f / (1.f + __expf(-zz));
            const float dz = dh[n * Hb + j] * (s * (1.f + zz * (1.f - s)));
            const float r = rstd ? rstd[n] : 1.f;
            a_db += dz;
            a_g0 += r * (mean ? mean[n] : 0.f) * dz;
            v = r * dz;
        }
        tile[ty + 8 * i][tx] = v;
    }
    s_db[ty][tx] = a_db; s_g0[ty][tx] = a_g0;
    __syncthreads();
    if (ty == 0) {
        float d = 0.f, g = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { d += s_db[k][tx]; g += s_g0[k][tx]; }
        if (j0 + tx < Hb) { atomicAdd(db1 + j0 + tx, d); if (g0) atomicAdd(g0 + j0 + tx, g); }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = j0 + ty + 8 * i;
        const int64_t n = n0 + tx;
        if (j < Hb && n < N) dzsT[(int64_t)j * tstride + n] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
}

// out[c] += Σ_r src[r, c]
template <typename Ti>
__global__ void __launch_bounds__(256)
colsum_kernel(const Ti* __restrict__ src, int64_t R, int C, int64_t sstride, float* __restrict__ out, int rows_per_cta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < R ? r0 + rows_per_cta : R;
    float a = 0.f;
    for (int64_t r = r0; r < r1; ++r) a += to_f32(src[r * sstride + c]);
    atomicAdd(out + c, a);
}

__global__ void __launch_bounds__(128)
wgrad_finish_kernel(const float* __restrict__ G, int64_t gstride, const float* __restrict__ w1, int64_t wstride,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ g0,
                    const float* __restrict__ db1, int Hb, int V, int rows_per_cta, float* __restrict__ dw1, int64_t dstride,
                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int j0 = blockIdx.y * rows_per_cta;
    const int j1 = min(j0 + rows_per_cta, Hb);
    const float gm = gamma[v], bt = beta[v];
    float ag = 0.f, ab = 0.f;
    for (int j = j0; j < j1; ++j) {
        const float d = G[(int64_t)j * gstride + v] - g0[j];
        const float w = w1[(int64_t)j * wstride + v];
        dw1[(int64_t)j * dstride + v] = fmaf(gm, d, bt * db1[j]);
        ag = fmaf(w, d, ag);
        ab = fmaf(w, db1[j], ab);
    }
    atomicAdd(dgamma + v, ag);
    atomicAdd(dbeta + v, ab);
}

// Softmax-attention backward, score gradient: dS[r,v] = P[r,v]·(dP[r,v] − δ_r) with δ_r = Σ_v P[r,v]·dP[r,v] = dZ[r,:]·Z[r,:]
// (Z = P·V, dP = dZ·Vᵀ), so δ needs only the d-wide output rows.  One CTA per row; bf16 K-major operand of dQ = dS·K.
__global__ void __launch_bounds__(256)
attn_ds_kernel(const __nv_bfloat16* __restrict__ P, int64_t pstride, const float* __restrict__ dP, int64_t dstride,
               const float* __restrict__ dZ, const float* __restrict__ Z, int64_t zstride, int d, int64_t rows, int V,
               __nv_bfloat16* __restrict__ dS, int64_t sstride) {
    __shared__ float red[8];
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        float part = 0.f;
        for (int c = threadIdx.x; c < d; c += blockDim.x) part = fmaf(dZ[r * zstride + c], Z[r * zstride + c], part);
        const float delta = block_sum_f(part, red);
        const __nv_bfloat16* p = P + r * pstride;
        const float* g = dP + r * dstride;
        __nv_bfloat16* o = dS + r * sstride;
        const int bd = blockDim.x;
        int c = threadIdx.x;
        for (; c + 3 * bd < V; c += 4 * bd) {
            const float p0 = __bfloat162float(p[c]), p1 = __bfloat162float(p[c + bd]), p2 = __bfloat162float(p[c + 2 * bd]),
                        p3 = __bfloat162float(p[c + 3 * bd]);
            const float g0 = g[c], g1 = g[c + bd], g2 = g[c + 2 * bd], g3 = g[c + 3 * bd];
            o[c] = __float2bfloat16_rn(p0 * (g0 - delta)); o[c + bd] = __float2bfloat16_rn(p1 * (g1 - delta));
            o[c + 2 * bd] = __float2bfloat16_rn(p2 * (g2 - delta)); o[c + 3 * bd] = __float2bfloat16_rn(p3 * (g3 - delta));
        }
        for (; c < V; c += bd) o[c] = __float2bfloat16_rn(__bfloat162float(p[c]) * (g[c] - delta));
        for (int64_t k = V + threadIdx.x; k < sstride; k += bd) o[k] = __float2bfloat16_rn(0.f);
        __syncthreads();                                       // `red` is reused by the next row
    }
}


// dst[i] = cast(scale * src[i]) over a flat buffer (gradient wire format: fp32 -> bf16 before the all-reduce, bf16 -> fp32
// with the 1/world averaging factor after it).  8 elements per thread, 128-bit accesses on the aligned body.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
flat_scale_cast_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int64_t n, float scale) {
    const int64_t n8 = n / 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        float v[8];
        if constexpr (sizeof(TI) == 4) {
            const uint4 a = ld_stream_u4(reinterpret_cast<const uint4*>(src) + 2 * i), b = ld_stream_u4(reinterpret_cast<const uint4*>(src) + 2 * i + 1);
            v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
            v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
        } else {
            unpack16(ld_stream_u4(reinterpret_cast<const uint4*>(src) + i), v, __nv_bfloat16());
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= scale;
        if constexpr (sizeof(TO) == 4) {
            reinterpret_cast<uint4*>(dst)[2 * i] = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
            reinterpret_cast<uint4*>(dst)[2 * i + 1] = make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
        } else {
            reinterpret_cast<uint4*>(dst)[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
    }
    for (int64_t i = n8 * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = from_f32<TO>(to_f32(src[i]) * scale);
}

}  // namespace tasu

using namespace tasu;

extern "C" int tasu_attn_score_grad(const void* P_bf16, int64_t p_stride, const float* dP, int64_t dp_stride,
                                    const float* dZ, const float* Z, int64_t z_stride, int d, int64_t rows, int V,
                                    void* dS_bf16, int64_t ds_stride, void* stream) {
    TASU_CHECK_ARG(rows >= 0 && V > 0 && d > 0 && p_stride >= V && dp_stride >= V && ds_stride >= V && z_stride >= d, "shape");
    if (rows == 0) return TASU_OK;
    TASU_CHECK_ARG(P_bf16 && dP && dZ && Z && dS_bf16, "null pointer");
    int64_t g = (int64_t)tasu::sm_count() * 8;
    if (g > rows) g = rows;
    attn_ds_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)P_bf16, p_stride, dP, dp_stride, dZ, Z,
                                                                  z_stride, d, rows, V, (__nv_bfloat16*)dS_bf16, ds_stride);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_transpose_cast(const void* src, int src_dtype, int64_t rows, int64_t cols, int64_t src_stride,
                                   const float* row_scale, void* dst_bf16, int64_t dst_stride, void* stream) {
    TASU_CHECK_ARG(rows >= 0 && cols >= 0, "rows, cols >= 0");
    TASU_CHECK_ARG(src_dtype == TASU_F32 || src_dtype == TASU_BF16, "src_dtype");
    TASU_CHECK_ARG(src_stride >= cols && dst_stride >= rows, "stride too small");
    if (rows == 0 || cols == 0) return TASU_OK;
    TASU_CHECK_ARG(src && dst_bf16, "null pointer");
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    TASU_CHECK_ARG(grid.y <= 65535u, "too many rows for one launch (max 2 097 120)");
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == TASU_F32)
        transpose_cast_kernel<float><<<grid, 256, 0, st>>>((const float*)src, rows, cols, src_stride, row_scale, (__nv_bfloat16*)dst_bf16, dst_stride);
    else
        transpose_cast_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, rows, cols, src_stride, row_scale, (__nv_bfloat16*)dst_bf16, dst_stride);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_silu_fwd(const float* z, int64_t n, void* h_bf16, void* stream) {
    TASU_CHECK_ARG(n >= 0, "n >= 0");
    if (n == 0) return TASU_OK;
    TASU_CHECK_ARG(z && h_bf16, "null pointer");
    TASU_CHECK_ARG(((uintptr_t)z % 16 == 0) && ((uintptr_t)h_bf16 % 8 == 0), "alignment");
    const int64_t threads = (n + 3) / 4;
    silu_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, n, (__nv_bfloat16*)h_bf16);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_silu_bwd(const float* dh, const float* z, int64_t N, int Hb, const float* row_rstd,
                             const float* row_mean, void* dzsT_bf16, int64_t t_stride, float* db1, float* g0,
                             void* stream) {
    TASU_CHECK_ARG(N >= 0 && Hb > 0 && t_stride >= N, "shape");
    TASU_CHECK_ARG(db1 != nullptr, "null db1");
    cudaStream_t st = (cudaStream_t)stream;
    TASU_CHECK_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * Hb, st));
    if (g0) TASU_CHECK_CUDA(cudaMemsetAsync(g0, 0, sizeof(float) * Hb, st));
    if (N == 0) return TASU_OK;
    TASU_CHECK_ARG(dh && z && dzsT_bf16, "null pointer");
    dim3 grid((unsigned)((Hb + 31) / 32), (unsigned)((N + 31) / 32));
    TASU_CHECK_ARG(grid.y <= 65535u, "too many rows for one launch");
    silu_bwd_kernel<<<grid, 256, 0, st>>>(dh, z, N, Hb, row_rstd, row_mean, (__nv_bfloat16*)dzsT_bf16, t_stride, db1, g0);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_colsum(const void* src, int src_dtype, int64_t rows, int cols, int64_t src_stride, float* out,
                           void* stream) {
    TASU_CHECK_ARG(rows >= 0 && cols > 0 && src_stride >= cols, "shape");
    TASU_CHECK_ARG(src_dtype == TASU_F32 || src_dtype == TASU_BF16, "src_dtype");
    TASU_CHECK_ARG(out != nullptr, "null out");
    cudaStream_t st = (cudaStream_t)stream;
    TASU_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
    if (rows == 0) return TASU_OK;
    TASU_CHECK_ARG(src != nullptr, "null src");
    const int rpc = 128;
    dim3 grid((unsigned)((cols + 255) / 256), (unsigned)((rows + rpc - 1) / rpc));
    TASU_CHECK_ARG(grid.y <= 65535u, "too many rows for one launch");
    if (src_dtype == TASU_F32) colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)src, rows, cols, src_stride, out, rpc);
    else colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, rows, cols, src_stride, out, rpc);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_linear_silu_wgrad_finish(const float* G, int64_t g_stride, const float* w1, int64_t w1_stride,
                                             const float* gamma, const float* beta, const float* g0, const float* db1,
                                             int Hb, int V,
                                             float* dw1, int64_t dw1_stride, float* dgamma, float* dbeta, void* stream) {
    TASU_CHECK_ARG(Hb > 0 && V > 0, "shape");
    TASU_CHECK_ARG(G && w1 && gamma && beta && g0 && db1 && dw1 && dgamma && dbeta, "null pointer");
    TASU_CHECK_ARG(g_stride >= V && w1_stride >= V && dw1_stride >= V, "stride too small");
    cudaStream_t st = (cudaStream_t)stream;
    TASU_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * V, st));
    TASU_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * V, st));
    const int rpc = 128;
    dim3 grid((unsigned)((V + 127) / 128), (unsigned)((Hb + rpc - 1) / rpc));
    wgrad_finish_kernel<<<grid, 128, 0, st>>>(G, g_stride, w1, w1_stride, gamma, beta, g0, db1, Hb, V, rpc, dw1, dw1_stride, dgamma, dbeta);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

// ---------------------------------------------------------------------------------------------
// fp32-accurate mode on the bf16 tensor cores: x = h1 + h2 + h3 (three bf16 terms, 24 mantissa bits);
// A·B ≈ h1h1 + h1h2 + h2h1 + h1h3 + h2h2 + h3h1 is ONE GEMM over a 6x longer K when the operands are laid
// out as  A' = [h1|h1|h2|h1|h2|h3]  and  B' = [h1|h2|h1|h3|h2|h1]  (blocks of pad64(K), zero padded).
namespace tasu {
template <typename Ti>
__global__ void __launch_bounds__(256)
split3_kernel(const Ti* __restrict__ src, int64_t rows, int K, int64_t sstride, const float* __restrict__ col_scale,
              int pattern, __nv_bfloat16* __restrict__ dst, int64_t dstride, int Kp, float* __restrict__ ln_mean,
              float* __restrict__ ln_rstd, float eps, float* __restrict__ row_sum) {
    __shared__ float red[8];
    // block order of (h1,h2,h3) per pattern: A = 0,0,1,0,1,2   B = 0,1,0,2,1,0
    const int pa[6] = {0, 0, 1, 0, 1, 2}, pb[6] = {0, 1, 0, 2, 1, 0};
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const Ti* s = src + r * sstride;
        __nv_bfloat16* d = dst + r * dstride;
        float acc_s = 0.f, acc_q = 0.f;
        for (int c = threadIdx.x; c < Kp; c += blockDim.x) {
            float x = 0.f;
            if (c < K) { x = to_f32(s[c]); if (col_scale) x *= col_scale[c]; }
            acc_s += x; acc_q += x * x;
            __nv_bfloat16 h[3];
            h[0] = __float2bfloat16_rn(x);
            const float r1 = x - __bfloat162float(h[0]);
            h[1] = __float2bfloat16_rn(r1);
            h[2] = __float2bfloat16_rn(r1 - __bfloat162float(h[1]));
#pragma unroll
            for (int b = 0; b < 6; ++b) d[(int64_t)b * Kp + c] = h[pattern == 0 ? pa[b] : pb[b]];
        }
        if (ln_mean != nullptr || row_sum != nullptr) {
            acc_s = block_sum_f(acc_s, red);
            acc_q = block_sum_f(acc_q, red);
            if (threadIdx.x == 0) {
                if (row_sum) row_sum[r] = acc_s;
                if (ln_mean) {
                    const float mean = acc_s / (float)K;
                    float var = acc_q / (float)K - mean * mean;
                    var = var < 0.f ? 0.f : var;
                    ln_mean[r] = mean;
                    ln_rstd[r] = rsqrtf(var + eps);
                }
            }
        }
    }
}
}  // namespace tasu

extern "C" int tasu_split_bf16x3(const void* src, int src_dtype, int64_t rows, int K, int64_t src_stride,
                                 const float* col_scale, int pattern, void* dst_bf16, int64_t dst_stride,
                                 float* ln_mean, float* ln_rstd, float ln_eps, float* row_sum, void* stream) {
    TASU_CHECK_ARG(rows >= 0 && K > 0 && src_stride >= K, "shape");
    TASU_CHECK_ARG(src_dtype == TASU_F32 || src_dtype == TASU_BF16, "src_dtype");
    TASU_CHECK_ARG(pattern == 0 || pattern == 1, "pattern: 0 = A (activations), 1 = B (weights)");
    TASU_CHECK_ARG((ln_mean == nullptr) == (ln_rstd == nullptr), "ln stats come in pairs");
    const int Kp = (K + 63) / 64 * 64;
    TASU_CHECK_ARG(dst_stride >= 6 * (int64_t)Kp, "dst_stride must hold 6 blocks of pad64(K)");
    if (rows == 0) return TASU_OK;
    TASU_CHECK_ARG(src && dst_bf16, "null pointer");
    int64_t g = (int64_t)tasu::sm_count() * 8;
    if (g > rows) g = rows;
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == TASU_F32)
        tasu::split3_kernel<float><<<(unsigned)g, 256, 0, st>>>((const float*)src, rows, K, src_stride, col_scale, pattern,
                                                               (__nv_bfloat16*)dst_bf16, dst_stride, Kp, ln_mean, ln_rstd, ln_eps, row_sum);
    else
        tasu::split3_kernel<__nv_bfloat16><<<(unsigned)g, 256, 0, st>>>((const __nv_bfloat16*)src, rows, K, src_stride, col_scale, pattern,
                                                                       (__nv_bfloat16*)dst_bf16, dst_stride, Kp, ln_mean, ln_rstd, ln_eps, row_sum);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

// Split-K combine for the fp32-accurate mode: out = epilogue(sum_p parts[p]) with round-to-nearest fp32 adds
// (the tensor core's own accumulation truncates; short K slices keep that bias below 1e-5).
namespace tasu {
template <typename To>
__global__ void __launch_bounds__(256)
sum_epilogue_kernel(const float* __restrict__ parts, int P, int64_t part_stride, int M, int N, int64_t ldp, int epi,
                    const float* __restrict__ bias, const float* __restrict__ rstd, const float* __restrict__ mean,
                    const float* __restrict__ colsum, To* __restrict__ out, int64_t ldo) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    const int m = (int)(idx / N), n = (int)(idx % N);
    float x = 0.f;
    for (int p = 0; p < P; ++p) x += parts[(int64_t)p * part_stride + (int64_t)m * ldp + n];
    if (epi == TASU_EPI_LNFOLD_SILU || epi == TASU_EPI_LNFOLD) {
        x = fmaf(rstd[m], x - mean[m] * colsum[n], bias[n]);
        if (epi == TASU_EPI_LNFOLD_SILU) x = x / (1.f + expf(-x));
    } else if (epi == TASU_EPI_SOFTMAX) {
        x = expf(x + bias[n] - mean[m]) * rstd[m];
    } else if (epi != TASU_EPI_NONE) {
        x += bias[n];
        if (epi == TASU_EPI_BIAS_SILU) x = x / (1.f + expf(-x));
        else if (epi == TASU_EPI_BIAS_RELU) x = fmaxf(x, 0.f);
    }
    out[(int64_t)m * ldo + n] = from_f32<To>(x);
}
}  // namespace tasu

extern "C" int tasu_sum_epilogue(const float* parts, int n_parts, int64_t part_stride, int M, int N, int64_t ldp,
                                 int epilogue, const float* bias, const float* row_rstd, const float* row_mean,
                                 const float* colsum, void* out, int out_dtype, int64_t ldo, void* stream) {
    TASU_CHECK_ARG(n_parts > 0 && M >= 0 && N > 0 && ldp >= N && ldo >= N, "shape");
    TASU_CHECK_ARG(out_dtype == TASU_F32 || out_dtype == TASU_BF16, "out_dtype");
    TASU_CHECK_ARG(epilogue >= TASU_EPI_NONE && epilogue <= TASU_EPI_SOFTMAX, "epilogue");
    if (M == 0) return TASU_OK;
    TASU_CHECK_ARG(parts && out, "null pointer");
    TASU_CHECK_ARG(epilogue == TASU_EPI_NONE || bias, "bias required");
    const int64_t n = (int64_t)M * N;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == TASU_F32)
        tasu::sum_epilogue_kernel<float><<<grid, 256, 0, st>>>(parts, n_parts, part_stride, M, N, ldp, epilogue, bias, row_rstd, row_mean, colsum, (float*)out, ldo);
    else
        tasu::sum_epilogue_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(parts, n_parts, part_stride, M, N, ldp, epilogue, bias, row_rstd, row_mean, colsum, (__nv_bfloat16*)out, ldo);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_flat_scale_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, float scale, void* stream) {
    TASU_CHECK_ARG(n >= 0, "n >= 0");
    TASU_CHECK_ARG((src_dtype == TASU_F32 || src_dtype == TASU_BF16) && (dst_dtype == TASU_F32 || dst_dtype == TASU_BF16), "dtype");
    if (n == 0) return TASU_OK;
    TASU_CHECK_ARG(src && dst, "null pointer");
    TASU_CHECK_ARG((uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0, "16-byte aligned buffers");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t g = (n / 8 + 255) / 256, gmax = (int64_t)tasu::sm_count() * 16;
    if (g > gmax) g = gmax;
    if (g < 1) g = 1;
    if (src_dtype == TASU_F32 && dst_dtype == TASU_BF16) flat_scale_cast_kernel<float, __nv_bfloat16><<<(unsigned)g, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n, scale);
    else if (src_dtype == TASU_BF16 && dst_dtype == TASU_F32) flat_scale_cast_kernel<__nv_bfloat16, float><<<(unsigned)g, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n, scale);
    else if (src_dtype == TASU_F32) flat_scale_cast_kernel<float, float><<<(unsigned)g, 256, 0, st>>>((const float*)src, (float*)dst, n, scale);
    else flat_scale_cast_kernel<__nv_bfloat16, __nv_bfloat16><<<(unsigned)g, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n, scale);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
