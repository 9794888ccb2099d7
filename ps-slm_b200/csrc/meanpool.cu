// Step 2c: segmented mean-pool of the kept candidates — Multitask/model/ps-slm.py:275-287
// (blank frame kept as is / non-blank run averaged), :297 (compaction), :303-314 (zero pad).
// One CTA per output row; the row (D up to ~152k elements) is streamed with 128-bit loads when
// every source row is 16-byte aligned (always true for buffers this library owns), otherwise
// with coalesced scalar loads (reference tensors with the odd 25055-element pitch).
// Optional fusions: softmax of logits rows on the fly (so the [N_in, V] posterior is never
// materialised) and LayerNorm statistics of the pooled row for the folded projector GEMM.
#include "common.cuh"

namespace tasu {

struct PoolArgs {
    const void* feats;
    int B, T, D;
    int64_t bstride, rstride;
    const float* smax;
    const float* ssum;
    const int32_t* seg_start;
    const int32_t* seg_len;
    const int32_t* row_off;
    const int32_t* seg_src;
    int layout;
    int64_t max_len, max_rows;
    void* out;
    int64_t ostride;
    float* ln_mean;
    float* ln_rstd;
    float eps;
};

// fp32 inputs are the accuracy path (1e-5 of the fp32 reference): full-precision expf; bf16 inputs keep the fast one
template <typename Tin> __device__ __forceinline__ float pool_exp(float x) { return sizeof(Tin) == 4 ? expf(x) : __expf(x); }

// locate packed row r: largest b with row_off[b] <= r (row_off has B+1 entries)
__device__ __forceinline__ int find_utt(const int32_t* __restrict__ row_off, int B, int r) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (row_off[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
}

template <typename Tin, typename Tout, bool kSoftmax, bool kVec>
__global__ void __launch_bounds__(256)
meanpool_kernel(PoolArgs a) {
    __shared__ float red[2][8];
    __shared__ int s_info[3];
    constexpr int VI = Vec16<Tin>::N;
    const Tin* feats = reinterpret_cast<const Tin*>(a.feats);
    Tout* out = reinterpret_cast<Tout*>(a.out);
    const int n_out = a.row_off[a.B];
    const int64_t total_rows = a.layout == 0 ? (int64_t)n_out : (int64_t)a.B * a.max_len;
    const int64_t limit = total_rows < a.max_rows ? total_rows : a.max_rows;

    for (int64_t r = blockIdx.x; r < limit; r += gridDim.x) {
        // ---- which candidate feeds this output row
        if (threadIdx.x == 0) {
            int b, j;
            if (a.layout == 0) { b = find_utt(a.row_off, a.B, (int)r); j = (int)r - a.row_off[b]; }
            else { b = (int)(r / a.max_len); j = (int)(r % a.max_len); }
            const int m_b = a.row_off[b + 1] - a.row_off[b];
            if (j < m_b) {
                s_info[0] = b;
                s_info[1] = a.seg_start[(int64_t)b * a.T + j];
                s_info[2] = a.seg_len[(int64_t)b * a.T + j];
            } else {
                s_info[0] = b; s_info[1] = 0; s_info[2] = 0;      // padded row → zeros
            }
        }
        __syncthreads();
        const int b = s_info[0], t0 = s_info[1], n = s_info[2];
        Tout* orow = out + r * a.ostride;
        const Tin* src = a.seg_src != nullptr ? feats + (int64_t)a.seg_src[r] * a.rstride
                                              : feats + (int64_t)b * a.bstride + (int64_t)t0 * a.rstride;
        float acc_s = 0.f, acc_q = 0.f;

        if (kVec) {
            // D-chunks of VI elements; output vector width follows the input chunk
            const int nchunk = a.D / VI;
            constexpr int U = 4;                 // independent 16-byte loads in flight per thread and frame
            for (int c0 = threadIdx.x; c0 < nchunk; c0 += U * blockDim.x) {
                float v[U][VI];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int e = 0; e < VI; ++e) v[u][e] = 0.f;
                for (int f = 0; f < n; ++f) {
                    const uint4* srow = reinterpret_cast<const uint4*>(src + (int64_t)f * a.rstride);
                    uint4 q[U];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (c0 + u * (int)blockDim.x < nchunk) q[u] = ld_stream_u4(srow + c0 + u * blockDim.x);
                    float mx = 0.f, is = 1.f;
                    if (kSoftmax) {
                        const int64_t fr = a.seg_src != nullptr ? (int64_t)a.seg_src[r] + f : (int64_t)b * a.T + t0 + f;
                        mx = a.smax[fr];
                        is = 1.f / a.ssum[fr];
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (c0 + u * (int)blockDim.x < nchunk) {
                            float x[VI];
                            unpack16(q[u], x, Tin());
#pragma unroll
                            for (int e = 0; e < VI; ++e) v[u][e] += kSoftmax ? pool_exp<Tin>(x[e] - mx) * is : x[e];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + u * blockDim.x;
                    if (c >= nchunk) break;
#pragma unroll
                    for (int e = 0; e < VI; ++e) {
                        if (n > 1) v[u][e] = v[u][e] / (float)n;
                        acc_s += v[u][e];
                        acc_q += v[u][e] * v[u][e];
                    }
                    Tout* o = orow + (int64_t)c * VI;
                    if constexpr (sizeof(Tout) == 4) {
#pragma unroll
                        for (int e = 0; e < VI; e += 4)
                            st_stream_u4(o + e, make_uint4(__float_as_uint(v[u][e]), __float_as_uint(v[u][e + 1]),
                                                           __float_as_uint(v[u][e + 2]), __float_as_uint(v[u][e + 3])));
                    } else if constexpr (VI == 8) {
                        st_stream_u4(o, make_uint4(pack_bf16x2(v[u][0], v[u][1]), pack_bf16x2(v[u][2], v[u][3]),
                                                   pack_bf16x2(v[u][4], v[u][5]), pack_bf16x2(v[u][6], v[u][7])));
                    } else {   // 4 fp32 in → 4 bf16 out (8 bytes)
                        *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(v[u][0], v[u][1]), pack_bf16x2(v[u][2], v[u][3]));
                    }
                }
            }
            // scalar remainder D % VI
            for (int d = nchunk * VI + threadIdx.x; d < a.D; d += blockDim.x) {
                float v = 0.f;
                for (int f = 0; f < n; ++f) {
                    float x = to_f32(src[(int64_t)f * a.rstride + d]);
                    if (kSoftmax) {
                        const int64_t fr = a.seg_src != nullptr ? (int64_t)a.seg_src[r] + f : (int64_t)b * a.T + t0 + f;
                        x = pool_exp<Tin>(x - a.smax[fr]) / a.ssum[fr];
                    }
                    v += x;
                }
                if (n > 1) v = v / (float)n;
                acc_s += v; acc_q += v * v;
                orow[d] = from_f32<Tout>(v);
            }
        } else {
            // rows with an odd pitch (reference tensors): coalesced 4-byte accesses, 4 columns in flight per thread
            const int bd = blockDim.x;
            int d = threadIdx.x;
            for (; d + 3 * bd < a.D; d += 4 * bd) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                for (int f = 0; f < n; ++f) {
                    const Tin* sr = src + (int64_t)f * a.rstride;
                    float x[4] = {to_f32(sr[d]), to_f32(sr[d + bd]), to_f32(sr[d + 2 * bd]), to_f32(sr[d + 3 * bd])};
                    if (kSoftmax) {
                        const int64_t fr = a.seg_src != nullptr ? (int64_t)a.seg_src[r] + f : (int64_t)b * a.T + t0 + f;
                        const float mx = a.smax[fr], is = 1.f / a.ssum[fr];
#pragma unroll
                        for (int e = 0; e < 4; ++e) x[e] = pool_exp<Tin>(x[e] - mx) * is;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] += x[e];
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (n > 1) v[e] = v[e] / (float)n;
                    acc_s += v[e]; acc_q += v[e] * v[e];
                    orow[d + e * bd] = from_f32<Tout>(v[e]);
                }
            }
            for (; d < a.D; d += bd) {
                float v = 0.f;
                for (int f = 0; f < n; ++f) {
                    float x = to_f32(src[(int64_t)f * a.rstride + d]);
                    if (kSoftmax) {
                        const int64_t fr = a.seg_src != nullptr ? (int64_t)a.seg_src[r] + f : (int64_t)b * a.T + t0 + f;
                        x = pool_exp<Tin>(x - a.smax[fr]) / a.ssum[fr];
                    }
                    v += x;
                }
                if (n > 1) v = v / (float)n;
                acc_s += v; acc_q += v * v;
                orow[d] = from_f32<Tout>(v);
            }
        }
        if (a.ln_mean != nullptr) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            acc_s = warp_sum(acc_s);
            acc_q = warp_sum(acc_q);
            if (lane == 0) { red[0][warp] = acc_s; red[1][warp] = acc_q; }
            __syncthreads();
            if (threadIdx.x == 0) {
                float s = 0.f, q = 0.f;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s += red[0][w]; q += red[1][w]; }
                const float mean = s / (float)a.D;
                float var = q / (float)a.D - mean * mean;
                var = var < 0.f ? 0.f : var;
                a.ln_mean[r] = mean;
                a.ln_rstd[r] = rsqrtf(var + a.eps);
            }
        }
        __syncthreads();   // s_info / red reuse
    }
}

}  // namespace tasu

using namespace tasu;

extern "C" int tasu_segment_meanpool(const void* feats, int in_dtype, int B, int T, int D,
                                     int64_t batch_stride, int64_t row_stride,
                                     const float* softmax_max, const float* softmax_sumexp,
                                     const int32_t* seg_start, const int32_t* seg_len, const int32_t* row_off,
                                     const int32_t* seg_src, int layout, int64_t max_len, int64_t max_rows,
                                     void* out, int out_dtype, int64_t out_row_stride,
                                     float* ln_mean, float* ln_rstd, float ln_eps, void* stream) {
    TASU_CHECK_ARG(B >= 0 && T >= 0 && D > 0, "B,T >= 0, D > 0");
    TASU_CHECK_ARG(in_dtype == TASU_F32 || in_dtype == TASU_BF16, "in_dtype");
    TASU_CHECK_ARG(out_dtype == TASU_F32 || out_dtype == TASU_BF16, "out_dtype");
    TASU_CHECK_ARG(layout == 0 || layout == 1, "layout");
    TASU_CHECK_ARG((softmax_max == nullptr) == (softmax_sumexp == nullptr), "softmax stats come in pairs");
    TASU_CHECK_ARG((ln_mean == nullptr) == (ln_rstd == nullptr), "ln stats come in pairs");
    TASU_CHECK_ARG(out_row_stride >= D, "out_row_stride < D");
    TASU_CHECK_ARG(seg_src == nullptr || layout == 0, "compact source needs the packed layout");
    if (B == 0 || max_rows <= 0 || (layout == 1 && max_len <= 0)) return TASU_OK;
    TASU_CHECK_ARG(feats && seg_start && seg_len && row_off && out, "null pointer");
    const int isz = in_dtype == TASU_F32 ? 4 : 2, osz = out_dtype == TASU_F32 ? 4 : 2;
    const int vi = 16 / isz;
    // vector path: every source row and every output row start on a 16-byte (resp. chunk) boundary
    const bool vec = ((uintptr_t)feats % 16 == 0) && ((batch_stride * isz) % 16 == 0) && ((row_stride * isz) % 16 == 0) &&
                     ((uintptr_t)out % 16 == 0) && ((out_row_stride * osz) % 16 == 0) && (D >= vi);
    PoolArgs a{feats, B, T, D, batch_stride, row_stride, softmax_max, softmax_sumexp, seg_start, seg_len, row_off,
               seg_src, layout, max_len, max_rows, out, out_row_stride, ln_mean, ln_rstd, ln_eps};
    int64_t rows_cap = layout == 0 ? max_rows : (int64_t)B * max_len;
    if (rows_cap > max_rows) rows_cap = max_rows;
    cudaStream_t st = (cudaStream_t)stream;
    const bool sm = softmax_max != nullptr;
#define LAUNCH1(K) K<<<persistent_grid(K, 256, rows_cap), 256, 0, st>>>(a)
#define LAUNCH(TI, TO)                                                                              \
    do {                                                                                            \
        if (sm) { if (vec) LAUNCH1((meanpool_kernel<TI, TO, true, true>));                          \
                  else     LAUNCH1((meanpool_kernel<TI, TO, true, false>)); }                       \
        else    { if (vec) LAUNCH1((meanpool_kernel<TI, TO, false, true>));                         \
                  else     LAUNCH1((meanpool_kernel<TI, TO, false, false>)); }                      \
    } while (0)
    if (in_dtype == TASU_F32 && out_dtype == TASU_F32) LAUNCH(float, float);
    else if (in_dtype == TASU_F32 && out_dtype == TASU_BF16) LAUNCH(float, __nv_bfloat16);
    else if (in_dtype == TASU_BF16 && out_dtype == TASU_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
    else LAUNCH(__nv_bfloat16, float);
#undef LAUNCH
#undef LAUNCH1
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
