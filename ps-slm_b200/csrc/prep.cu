// Operand producers for the projector GEMMs and the simulated-posterior constructor.
//  * cast_rows        generic [rows, cols] dtype conversion (+ LayerNorm statistics per row):
//                     the A-operand producer for EncoderProjectorLinearSiLU (projector.py:150)
//  * fold_layernorm   W1g = bf16(W1 * gamma), colsum, dbias: folds nn.LayerNorm(25055)
//                     (projector.py:139) into nn.Linear(25055, 2048) (projector.py:141)
//  * sim_posterior    rows (1-alpha)*onehot + alpha/V / hard blank / zero pad
//                     (ps-slm.py:346-358, :380-408) written straight in HBM.
#include "common.cuh"
#include <cooperative_groups.h>

namespace tasu {

template <typename Ti, typename To>
__global__ void __launch_bounds__(256)
cast_rows_kernel(const Ti* __restrict__ src, int64_t rows, int cols, int64_t sstride, To* __restrict__ dst,
                 int64_t dstride, float* __restrict__ ln_mean, float* __restrict__ ln_rstd, float eps) {
    __shared__ float red[2][8];
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const Ti* s = src + r * sstride;
        To* d = dst + r * dstride;
        float acc_s = 0.f, acc_q = 0.f;
        const int bd = blockDim.x;
        int c = threadIdx.x;
        for (; c + 3 * bd < cols; c += 4 * bd) {             // 4 independent coalesced loads in flight per thread
            const float v0 = to_f32(s[c]), v1 = to_f32(s[c + bd]), v2 = to_f32(s[c + 2 * bd]), v3 = to_f32(s[c + 3 * bd]);
            acc_s += (v0 + v1) + (v2 + v3);
            acc_q += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
            d[c] = from_f32<To>(v0); d[c + bd] = from_f32<To>(v1); d[c + 2 * bd] = from_f32<To>(v2); d[c + 3 * bd] = from_f32<To>(v3);
        }
        for (; c < cols; c += bd) {
            const float v = to_f32(s[c]);
            acc_s += v;
            acc_q += v * v;
            d[c] = from_f32<To>(v);
        }
        if (ln_mean != nullptr) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            acc_s = warp_sum(acc_s);
            acc_q = warp_sum(acc_q);
            __syncthreads();
            if (lane == 0) { red[0][warp] = acc_s; red[1][warp] = acc_q; }
            __syncthreads();
            if (threadIdx.x == 0) {
                float sm = 0.f, q = 0.f;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sm += red[0][w]; q += red[1][w]; }
                const float mean = sm / (float)cols;
                float var = q / (float)cols - mean * mean;
                var = var < 0.f ? 0.f : var;
                ln_mean[r] = mean;
                ln_rstd[r] = rsqrtf(var + eps);
            }
        }
    }
}

// fast path: fp32 → bf16, rows of 8-element multiples, 16-byte aligned: 2 x 128-bit loads → 1 x 128-bit store
__global__ void __launch_bounds__(256)
cast_f32_bf16_vec_kernel(const float* __restrict__ src, int64_t rows, int cols8, int64_t sstride, __nv_bfloat16* __restrict__ dst,
                         int64_t dstride) {
    const int64_t total = rows * cols8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols8;
        const int c = (int)(i % cols8);
        const uint4 a = ld_stream_u4(reinterpret_cast<const uint4*>(src + r * sstride) + 2 * c);
        const uint4 b = ld_stream_u4(reinterpret_cast<const uint4*>(src + r * sstride) + 2 * c + 1);
        const uint4 o = make_uint4(pack_bf16x2(__uint_as_float(a.x), __uint_as_float(a.y)), pack_bf16x2(__uint_as_float(a.z), __uint_as_float(a.w)),
                                   pack_bf16x2(__uint_as_float(b.x), __uint_as_float(b.y)), pack_bf16x2(__uint_as_float(b.z), __uint_as_float(b.w)));
        reinterpret_cast<uint4*>(dst + r * dstride)[c] = o;
    }
}

// fp32 → bf16 row cast that also emits the squared 2-norm of every (fp32) row: one warp per row.  The exact-decision
// error bound of the fused CTC head needs ||x_f|| per frame (csrc/refine.cu); taken here it costs no extra pass.
__global__ void __launch_bounds__(256)
cast_f32_bf16_sumsq_kernel(const float* __restrict__ src, int64_t rows, int cols, int64_t sstride, __nv_bfloat16* __restrict__ dst,
                           int64_t dstride, int64_t dcols, float* __restrict__ row_sumsq, int vec) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nwarp) {
        const float* s = src + r * sstride;
        __nv_bfloat16* d = dst + r * dstride;
        float q = 0.f;
        if (vec) {
            for (int c = lane; c < cols / 8; c += 32) {
                const uint4 a = ld_stream_u4(reinterpret_cast<const uint4*>(s) + 2 * c), b = ld_stream_u4(reinterpret_cast<const uint4*>(s) + 2 * c + 1);
                const float v[8] = {__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w),
                                    __uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w)};
#pragma unroll
                for (int e = 0; e < 8; ++e) q = fmaf(v[e], v[e], q);
                reinterpret_cast<uint4*>(d)[c] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
        } else {
            for (int c = lane; c < cols; c += 32) { const float v = s[c]; q = fmaf(v, v, q); d[c] = __float2bfloat16_rn(v); }
        }
        for (int64_t c = cols + lane; c < dcols; c += 32) d[c] = __float2bfloat16_rn(0.f);     // pad columns of the pitch
        q = warp_sum(q);
        if (lane == 0) row_sumsq[r] = q;
    }
}

// one CTA per output feature n (row of W1)
__global__ void __launch_bounds__(256)
fold_layernorm_kernel(const float* __restrict__ w1, int64_t wstride, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ b1, int K,
                      __nv_bfloat16* __restrict__ w1g, int64_t gstride, float* __restrict__ colsum,
                      float* __restrict__ dbias) {
    __shared__ float red[8];
    const int n = blockIdx.x;
    const float* w = w1 + (int64_t)n * wstride;
    __nv_bfloat16* g = w1g + (int64_t)n * gstride;
    float cs = 0.f, db = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float wv = w[k];
        const __nv_bfloat16 q = __float2bfloat16_rn(wv * gamma[k]);
        g[k] = q;
        cs += __bfloat162float(q);            // sum of the ROUNDED values: what the tensor core multiplies
        db += wv * beta[k];
    }
    // zero the pitch padding so a K-padded TMA box never sees garbage
    for (int64_t k = K + threadIdx.x; k < gstride; k += blockDim.x) g[k] = __float2bfloat16_rn(0.f);
    cs = block_sum_f(cs, red);
    db = block_sum_f(db, red);
    if (threadIdx.x == 0) {
        colsum[n] = cs;
        dbias[n] = db + (b1 ? b1[n] : 0.f);
    }
}

template <typename To>
__global__ void __launch_bounds__(256)
sim_rows_kernel(const int32_t* __restrict__ tok, const float* __restrict__ hot, const float* __restrict__ base,
                const int64_t* __restrict__ dst_row, int64_t n_rows, int V, To* __restrict__ out, int64_t ostride,
                float* __restrict__ ln_mean, float* __restrict__ ln_rstd, float eps) {
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int id = tok[r];
        const float bv = id < 0 ? 0.f : base[r];
        const float hv = id < 0 ? 0.f : hot[r];
        const int64_t dr = dst_row ? dst_row[r] : r;
        To* o = out + dr * ostride;
        const To bq = from_f32<To>(bv);
        for (int c = threadIdx.x; c < V; c += blockDim.x) o[c] = (c == id) ? from_f32<To>(hv) : bq;
        if (ln_mean != nullptr && threadIdx.x == 0) {
            // closed form: one element hv, V-1 elements bv (double: no cancellation issues)
            const double s = (id < 0) ? 0.0 : (double)hv + (double)(V - 1) * (double)bv;
            const double q = (id < 0) ? 0.0 : (double)hv * hv + (double)(V - 1) * (double)bv * bv;
            const double mean = s / V;
            double var = q / V - mean * mean;
            var = var < 0 ? 0 : var;
            ln_mean[dr] = (float)mean;
            ln_rstd[dr] = (float)(1.0 / sqrt(var + (double)eps));
        }
    }
}

// out[r, v] = bf16(exp(x[r, v] - row_max[r]) / row_sumexp[r]) for v < V, 0 for V <= v < ldo: the softmax of rows whose
// statistics are known (tasu_frame_stats), written as the K-major bf16 A operand of a following contraction
// (voca_trans: softmax(logits_no_blank) · embed_matrix, ps-slm.py:494-497).  One CTA per row, grid-stride.
template <typename Ti>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const Ti* __restrict__ x, int64_t xstride, int64_t rows, int V, const float* __restrict__ row_max,
                    const float* __restrict__ row_sumexp, __nv_bfloat16* __restrict__ out, int64_t ostride) {
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const Ti* s = x + r * xstride;
        __nv_bfloat16* d = out + r * ostride;
        const float m = row_max[r] * 1.4426950408889634f, inv = 1.f / row_sumexp[r];
        const int bd = blockDim.x;
        int c = threadIdx.x;
        for (; c + 3 * bd < V; c += 4 * bd) {                   // 4 independent coalesced loads in flight per thread
            const float v0 = to_f32(s[c]), v1 = to_f32(s[c + bd]), v2 = to_f32(s[c + 2 * bd]), v3 = to_f32(s[c + 3 * bd]);
            d[c] = __float2bfloat16_rn(exp2f(fmaf(v0, 1.4426950408889634f, -m)) * inv);
            d[c + bd] = __float2bfloat16_rn(exp2f(fmaf(v1, 1.4426950408889634f, -m)) * inv);
            d[c + 2 * bd] = __float2bfloat16_rn(exp2f(fmaf(v2, 1.4426950408889634f, -m)) * inv);
            d[c + 3 * bd] = __float2bfloat16_rn(exp2f(fmaf(v3, 1.4426950408889634f, -m)) * inv);
        }
        for (; c < V; c += bd) d[c] = __float2bfloat16_rn(exp2f(fmaf(to_f32(s[c]), 1.4426950408889634f, -m)) * inv);
        for (int64_t k = V + threadIdx.x; k < ostride; k += bd) d[k] = __float2bfloat16_rn(0.f);
    }
}

// one warp per kept candidate: copy its frames' encoder rows (K bf16 each) into the compact matrix.
// Row r < n_out holds the candidate's first frame, extra frames of multi-frame runs go to the tail.
__global__ void __launch_bounds__(256)
gather_kept_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int B, int T, int n_prefix, int K, int V,
                        const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_len,
                        const int32_t* __restrict__ seg_foff, const int32_t* __restrict__ row_off,
                        const int32_t* __restrict__ frame_off, const float* __restrict__ row_max,
                        const float* __restrict__ row_sumexp, const float* __restrict__ row_sumexp2,
                        int64_t max_rows, int64_t max_out, __nv_bfloat16* __restrict__ xg, int64_t ldg, float* __restrict__ g_max,
                        float* __restrict__ g_inv, int32_t* __restrict__ pk_len, int32_t* __restrict__ tail_src,
                        int32_t* __restrict__ multi_rows, int32_t* __restrict__ multi_count,
                        float* __restrict__ ln_mean, float* __restrict__ ln_rstd, float eps) {
    const int lane = threadIdx.x & 31;
    const int n_out = row_off[B];
    // capacities (max_out packed rows, max_rows compact rows) bound every write: an overflowing batch produces
    // an incomplete result that the host detects from the header and redoes with larger buffers
    const int n_loop = n_out < max_out ? n_out : (int)max_out;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const bool vec = (K % 8 == 0) && ((ldx % 8) == 0) && ((ldg % 8) == 0);
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_loop; r += warps) {
        int lo = 0, hi = B;                                    // utterance of packed row r
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (row_off[mid] <= r) lo = mid; else hi = mid; }
        const int b = lo, j = r - row_off[b];
        const int64_t pj = (int64_t)b * T + j;
        const int t0 = seg_start[pj], n = seg_len[pj];
        // extra frames of earlier candidates: (kept frames before) - (kept candidates before)
        const int tail0 = n_out + (frame_off[b] - row_off[b]) + (seg_foff[pj] - j);
        if (lane == 0) {
            pk_len[r] = n;
            tail_src[r] = tail0;
            if (n > 1 && multi_rows != nullptr) multi_rows[atomicAdd(multi_count, 1)] = r;   // work list of pool_tail
        }
        for (int f = 0; f < n; ++f) {
            const int64_t drow = f == 0 ? r : (int64_t)tail0 + f - 1;
            if (drow >= max_rows) break;
            const __nv_bfloat16* src = x + ((int64_t)b * (T + n_prefix) + n_prefix + t0 + f) * ldx;
            __nv_bfloat16* dst = xg + drow * ldg;
            if (vec) {
                for (int c = lane; c < K / 8; c += 32)
                    reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(src)[c];
            } else {
                for (int c = lane; c < K; c += 32) dst[c] = src[c];
            }
            if (lane == 0) {
                const int64_t fr = (int64_t)b * T + t0 + f;
                const float s = row_sumexp[fr];
                g_max[drow] = row_max[fr];
                g_inv[drow] = 1.f / s;
                if (f == 0 && ln_mean != nullptr) {
                    // single-frame row: mean p = 1/V, sum p^2 = s2/s^2 (multi-frame rows are overwritten by pool_tail)
                    const float mean = 1.f / (float)V;
                    const float q = row_sumexp2 ? row_sumexp2[fr] / (s * s) : 0.f;
                    float var = q / (float)V - mean * mean;
                    var = var < 0.f ? 0.f : var;
                    ln_mean[r] = mean;
                    ln_rstd[r] = rsqrtf(var + eps);
                }
            }
        }
    }
}

// multi-frame candidates only: probs[r] = mean over its frames, in place; LayerNorm statistics of the result.
// Work list = multi_rows[0, *multi_count) (any order) or, without it, every row with pk_len > 1.
__global__ void __launch_bounds__(256, 5)
pool_tail_kernel(__nv_bfloat16* __restrict__ probs, int64_t ld, int D, int64_t n_out, int64_t max_rows,
                 const int32_t* __restrict__ pk_len,
                 const int32_t* __restrict__ tail_src, const int32_t* __restrict__ multi_rows,
                 const int32_t* __restrict__ multi_count, float* __restrict__ ln_mean, float* __restrict__ ln_rstd, float eps) {
    __shared__ float red[8];
    const int64_t n_items = multi_rows != nullptr ? (int64_t)*multi_count : n_out;
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int64_t r = multi_rows != nullptr ? (int64_t)multi_rows[it] : it;
        const int n = pk_len[r];
        if (n <= 1) continue;                                  // CTA-uniform
        // a row whose frames do not all fit the compact matrix (capacity exceeded: the host redoes the batch with
        // larger buffers) is left alone — nothing is read or written beyond max_rows
        if (r >= max_rows || (int64_t)tail_src[r] + (n - 1) > max_rows) continue;
        __nv_bfloat16* row = probs + r * ld;
        const __nv_bfloat16* tail = probs + (int64_t)tail_src[r] * ld;
        const float inv = 1.f / (float)n;
        float acc_q = 0.f;
        const int nchunk = D / 8;
        const int bd = blockDim.x;
        for (int c = threadIdx.x; c < nchunk; c += 2 * bd) {
            const bool two = c + bd < nchunk;
            // issue every load of both chunks before the first add (runs are 2-3 frames almost always)
            uint4 h0 = ld_stream_u4(reinterpret_cast<const uint4*>(row) + c), h1 = h0;
            uint4 a0 = ld_stream_u4(reinterpret_cast<const uint4*>(tail) + c), a1 = a0, b0 = a0, b1 = a0;
            if (two) { h1 = ld_stream_u4(reinterpret_cast<const uint4*>(row) + c + bd);
                       a1 = ld_stream_u4(reinterpret_cast<const uint4*>(tail) + c + bd); }
            if (n > 2) { b0 = ld_stream_u4(reinterpret_cast<const uint4*>(tail + ld) + c);
                         if (two) b1 = ld_stream_u4(reinterpret_cast<const uint4*>(tail + ld) + c + bd); }
            float v0[8], v1[8], x[8];
            unpack16(h0, v0, __nv_bfloat16()); unpack16(h1, v1, __nv_bfloat16());
            unpack16(a0, x, __nv_bfloat16());
#pragma unroll
            for (int e = 0; e < 8; ++e) v0[e] += x[e];
            unpack16(a1, x, __nv_bfloat16());
#pragma unroll
            for (int e = 0; e < 8; ++e) v1[e] += x[e];
            if (n > 2) {
                unpack16(b0, x, __nv_bfloat16());
#pragma unroll
                for (int e = 0; e < 8; ++e) v0[e] += x[e];
                unpack16(b1, x, __nv_bfloat16());
#pragma unroll
                for (int e = 0; e < 8; ++e) v1[e] += x[e];
            }
            for (int f = 3; f < n; ++f) {                      // long runs (rare)
                unpack16(ld_stream_u4(reinterpret_cast<const uint4*>(tail + (int64_t)(f - 1) * ld) + c), x, __nv_bfloat16());
#pragma unroll
                for (int e = 0; e < 8; ++e) v0[e] += x[e];
                if (two) {
                    unpack16(ld_stream_u4(reinterpret_cast<const uint4*>(tail + (int64_t)(f - 1) * ld) + c + bd), x, __nv_bfloat16());
#pragma unroll
                    for (int e = 0; e < 8; ++e) v1[e] += x[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { v0[e] *= inv; acc_q += v0[e] * v0[e]; }
            reinterpret_cast<uint4*>(row)[c] = make_uint4(pack_bf16x2(v0[0], v0[1]), pack_bf16x2(v0[2], v0[3]),
                                                         pack_bf16x2(v0[4], v0[5]), pack_bf16x2(v0[6], v0[7]));
            if (two) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { v1[e] *= inv; acc_q += v1[e] * v1[e]; }
                reinterpret_cast<uint4*>(row)[c + bd] = make_uint4(pack_bf16x2(v1[0], v1[1]), pack_bf16x2(v1[2], v1[3]),
                                                                  pack_bf16x2(v1[4], v1[5]), pack_bf16x2(v1[6], v1[7]));
            }
        }
        for (int d = nchunk * 8 + threadIdx.x; d < D; d += blockDim.x) {
            float v = __bfloat162float(row[d]);
            for (int f = 1; f < n; ++f) v += __bfloat162float(tail[(int64_t)(f - 1) * ld + d]);
            v *= inv;
            acc_q += v * v;
            row[d] = __float2bfloat16_rn(v);
        }
        acc_q = block_sum_f(acc_q, red);
        if (threadIdx.x == 0 && ln_mean != nullptr) {
            const float mean = 1.f / (float)D;
            float var = acc_q / (float)D - mean * mean;
            var = var < 0.f ? 0.f : var;
            ln_mean[r] = mean;
            ln_rstd[r] = rsqrtf(var + eps);
        }
    }
}

// Natural-order index of the kept frames (fp32-accurate path): packed candidate r covers compact rows
// [seg_src[r], seg_src[r] + len) and frame_row[compact row] = its raw encoder row.  One thread per candidate.
__global__ void __launch_bounds__(256)
kept_frame_index_kernel(const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_len,
                        const int32_t* __restrict__ seg_foff, const int32_t* __restrict__ row_off,
                        const int32_t* __restrict__ frame_off, int B, int T, int n_prefix, int64_t max_rows, int64_t max_out,
                        int32_t* __restrict__ frame_row, int32_t* __restrict__ seg_src) {
    const int n_out = row_off[B];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_out || r >= max_out) return;
    int lo = 0, hi = B;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (row_off[mid] <= r) lo = mid; else hi = mid; }
    const int b = lo, j = r - row_off[b];
    const int64_t pj = (int64_t)b * T + j;
    const int t0 = seg_start[pj], n = seg_len[pj];
    const int c0 = frame_off[b] + seg_foff[pj];
    seg_src[r] = c0;
    for (int f = 0; f < n; ++f)
        if (c0 + f < max_rows) frame_row[c0 + f] = b * (T + n_prefix) + n_prefix + t0 + f;
}

// Content fingerprint of up to 8 buffers (weight-cache validation): 4096 32-bit words of every buffer — 32 evenly spaced
// windows of 128 consecutive words (coalesced 128-byte lines, at most 32 pages per buffer: 4096 scattered single words
// cost 22 us in sector requests and page walks, the windows a few) — each multiplied by an odd constant that depends on
// its sample index, summed modulo 2^64.  One CTA; thread 0 stores the result with a plain store, so `out` may live in
// pinned host memory.
struct FingerprintArgs { const uint32_t* ptr[8]; int64_t words[8]; int64_t stride[8]; int n; };
constexpr int kFpSamples = 4096;

// One thread-block cluster of 8 CTAs, CTA t takes buffer t (its 4096 samples as 8 independent loads per thread); the
// partial sums meet in the shared memory of CTA 0 (distributed shared memory), which stores the result.  One CTA for
// all buffers needed 20 us on its single SM; the cluster needs a few.
// Sample k of buffer t is word (k / 128) * stride[t] + k % 128 (stride = 128 for buffers of at most kFpSamples words:
// sample k = word k; else the host spreads the 32 windows evenly).
constexpr int kFpThreads = 512;
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(kFpThreads)
fingerprint_kernel(const FingerprintArgs a, unsigned long long* __restrict__ out) {
    __shared__ unsigned long long red[kFpThreads / 32];
    __shared__ unsigned long long part[8];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int t = (int)cluster.block_rank();
    constexpr int kPer = kFpSamples / kFpThreads;
    unsigned long long h = 0;
    if (t < a.n) {
        const uint32_t* __restrict__ ptr = a.ptr[t];
        const int64_t words = a.words[t], stride = a.stride[t];
        uint32_t v[kPer];
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int k = (int)threadIdx.x + kFpThreads * j;
            const int64_t idx = (int64_t)(k >> 7) * stride + (k & 127);
            v[j] = idx < words ? __ldg(ptr + idx) : 0u;
        }
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int k = (int)threadIdx.x + kFpThreads * j;
            const int i = t * kFpSamples + k;
            if ((int64_t)(k >> 7) * stride + (k & 127) < words)
                h += ((unsigned long long)v[j] + 0x9E3779B97F4A7C15ull) * (2ull * (unsigned long long)(i + 1) * 0xD6E8FEB86659FD93ull + 1ull);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int i = 0; i < kFpThreads / 32; ++i) s += red[i];
        *cluster.map_shared_rank(&part[t], 0) = s;             // CTA 0's part[t]
    }
    cluster.sync();
    if (t == 0 && threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int i = 0; i < 8; ++i) s += part[i];
        *out = s;
    }
}

}  // namespace tasu

using namespace tasu;

static unsigned row_grid(int64_t rows) {
    int64_t g = (int64_t)tasu::sm_count() * 8;
    if (g > rows) g = rows;
    return (unsigned)(g < 1 ? 1 : g);
}

extern "C" int tasu_cast_rows(const void* src, int src_dtype, int64_t rows, int cols, int64_t src_stride,
                              void* dst, int dst_dtype, int64_t dst_stride, float* ln_mean, float* ln_rstd,
                              float ln_eps, void* stream) {
    TASU_CHECK_ARG(rows >= 0 && cols > 0, "rows >= 0, cols > 0");
    TASU_CHECK_ARG(src_dtype == TASU_F32 || src_dtype == TASU_BF16, "src_dtype");
    TASU_CHECK_ARG(dst_dtype == TASU_F32 || dst_dtype == TASU_BF16, "dst_dtype");
    TASU_CHECK_ARG((ln_mean == nullptr) == (ln_rstd == nullptr), "ln stats come in pairs");
    if (rows == 0) return TASU_OK;
    TASU_CHECK_ARG(src && dst, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == TASU_F32 && dst_dtype == TASU_BF16 && ln_mean == nullptr && cols % 8 == 0 && src_stride % 4 == 0 &&
        dst_stride % 8 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
        const int64_t total = rows * (cols / 8);
        int64_t g = (total + 255) / 256, gmax = (int64_t)tasu::sm_count() * 16;
        if (g > gmax) g = gmax;
        cast_f32_bf16_vec_kernel<<<(unsigned)g, 256, 0, st>>>((const float*)src, rows, cols / 8, src_stride, (__nv_bfloat16*)dst, dst_stride);
        TASU_CHECK_LAUNCH();
        return TASU_OK;
    }
    const unsigned grid = row_grid(rows);
#define LAUNCH(TI, TO) cast_rows_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)src, rows, cols, src_stride, (TO*)dst, dst_stride, ln_mean, ln_rstd, ln_eps)
    if (src_dtype == TASU_F32 && dst_dtype == TASU_BF16) LAUNCH(float, __nv_bfloat16);
    else if (src_dtype == TASU_F32) LAUNCH(float, float);
    else if (dst_dtype == TASU_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
    else LAUNCH(__nv_bfloat16, float);
#undef LAUNCH
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_fold_layernorm(const float* w1, int64_t w1_stride, const float* gamma, const float* beta,
                                   const float* b1, int N, int K, void* w1g_bf16, int64_t w1g_stride,
                                   float* colsum, float* dbias, void* stream) {
    TASU_CHECK_ARG(N > 0 && K > 0, "N,K > 0");
    TASU_CHECK_ARG(w1 && gamma && beta && w1g_bf16 && colsum && dbias, "null pointer");
    TASU_CHECK_ARG(w1g_stride >= K && w1_stride >= K, "stride < K");
    fold_layernorm_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(w1, w1_stride, gamma, beta, b1, K,
                                                              (__nv_bfloat16*)w1g_bf16, w1g_stride, colsum, dbias);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_sim_posterior_rows(const int32_t* tok, const float* hot, const float* base,
                                       const int64_t* dst_row, int64_t n_rows, int V, void* out, int out_dtype,
                                       int64_t out_row_stride, float* ln_mean, float* ln_rstd, float ln_eps,
                                       void* stream) {
    TASU_CHECK_ARG(n_rows >= 0 && V > 0, "n_rows >= 0, V > 0");
    TASU_CHECK_ARG(out_dtype == TASU_F32 || out_dtype == TASU_BF16, "out_dtype");
    TASU_CHECK_ARG(out_row_stride >= V, "out_row_stride < V");
    TASU_CHECK_ARG((ln_mean == nullptr) == (ln_rstd == nullptr), "ln stats come in pairs");
    if (n_rows == 0) return TASU_OK;
    TASU_CHECK_ARG(tok && hot && base && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = row_grid(n_rows);
    if (out_dtype == TASU_F32)
        sim_rows_kernel<float><<<grid, 256, 0, st>>>(tok, hot, base, dst_row, n_rows, V, (float*)out, out_row_stride, ln_mean, ln_rstd, ln_eps);
    else
        sim_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(tok, hot, base, dst_row, n_rows, V, (__nv_bfloat16*)out, out_row_stride, ln_mean, ln_rstd, ln_eps);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_gather_kept_rows(const void* x_bf16, int64_t ldx, int B, int T, int n_prefix, int K, int V,
                                     const int32_t* seg_start, const int32_t* seg_len, const int32_t* seg_frame_off,
                                     const int32_t* row_off, const int32_t* frame_off, const float* row_max,
                                     const float* row_sumexp, const float* row_sumexp2, int64_t max_rows,
                                     int64_t max_out, void* xg_bf16, int64_t ldg, float* g_max, float* g_inv_sum, int32_t* pk_len,
                                     int32_t* tail_src, int32_t* multi_rows, int32_t* multi_count, float* ln_mean,
                                     float* ln_rstd, float ln_eps, void* stream) {
    TASU_CHECK_ARG(B >= 0 && T >= 0 && n_prefix >= 0 && K > 0 && V > 0 && ldx >= K && ldg >= K, "shape");
    TASU_CHECK_ARG((multi_rows == nullptr) == (multi_count == nullptr), "multi_rows / multi_count come in pairs");
    if (multi_count) TASU_CHECK_CUDA(cudaMemsetAsync(multi_count, 0, sizeof(int32_t), (cudaStream_t)stream));
    TASU_CHECK_ARG((ln_mean == nullptr) == (ln_rstd == nullptr), "ln stats come in pairs");
    if (B == 0 || max_rows <= 0 || max_out <= 0) return TASU_OK;
    TASU_CHECK_ARG(x_bf16 && seg_start && seg_len && seg_frame_off && row_off && frame_off && row_max && row_sumexp &&
                   xg_bf16 && g_max && g_inv_sum && pk_len && tail_src, "null pointer");
    TASU_CHECK_ARG(((uintptr_t)x_bf16 % 16 == 0) && ((uintptr_t)xg_bf16 % 16 == 0), "16-byte alignment");
    const unsigned grid = (unsigned)(tasu::sm_count() * 8);
    gather_kept_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x_bf16, ldx, B, T, n_prefix, K, V, seg_start, seg_len, seg_frame_off, row_off, frame_off,
        row_max, row_sumexp, row_sumexp2, max_rows, max_out, (__nv_bfloat16*)xg_bf16, ldg, g_max, g_inv_sum, pk_len, tail_src,
        multi_rows, multi_count, ln_mean, ln_rstd, ln_eps);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_pool_tail(void* probs_bf16, int64_t ld, int D, int64_t n_out, int64_t max_rows, const int32_t* pk_len,
                              const int32_t* tail_src, const int32_t* multi_rows, const int32_t* multi_count,
                              float* ln_mean, float* ln_rstd, float ln_eps, void* stream) {
    TASU_CHECK_ARG(D > 0 && ld >= D && n_out >= 0 && max_rows >= 0, "shape");
    TASU_CHECK_ARG((ln_mean == nullptr) == (ln_rstd == nullptr), "ln stats come in pairs");
    if (n_out == 0 || max_rows == 0) return TASU_OK;
    TASU_CHECK_ARG(probs_bf16 && pk_len && tail_src, "null pointer");
    TASU_CHECK_ARG(((uintptr_t)probs_bf16 % 16 == 0) && (ld % 8 == 0), "16-byte aligned rows");
    TASU_CHECK_ARG((multi_rows == nullptr) == (multi_count == nullptr), "multi_rows / multi_count come in pairs");
    // One CTA per potential work item (the live count is on the device; surplus CTAs exit at once): the hardware CTA
    // scheduler then balances the ~2 k multi-frame rows dynamically — a persistent grid of SMs x occupancy CTAs leaves
    // half the machine idle while the CTAs that drew 3 rows instead of 2 finish.
    const int64_t grid = n_out < 0x7fffffffLL ? n_out : 0x7fffffffLL;
    pool_tail_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)probs_bf16, ld, D, n_out, max_rows, pk_len,
                                                                         tail_src, multi_rows, multi_count, ln_mean,
                                                                         ln_rstd, ln_eps);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_softmax_rows(const void* x, int x_dtype, int64_t x_row_stride, int64_t rows, int V, const float* row_max,
                                 const float* row_sumexp, void* out_bf16, int64_t out_row_stride, void* stream) {
    TASU_CHECK_ARG(rows >= 0 && V > 0 && x_row_stride >= V && out_row_stride >= V, "shape");
    TASU_CHECK_ARG(x_dtype == TASU_F32 || x_dtype == TASU_BF16, "x_dtype");
    if (rows == 0) return TASU_OK;
    TASU_CHECK_ARG(x && row_max && row_sumexp && out_bf16, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = row_grid(rows);
    if (x_dtype == TASU_F32)
        softmax_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)x, x_row_stride, rows, V, row_max, row_sumexp,
                                                         (__nv_bfloat16*)out_bf16, out_row_stride);
    else
        softmax_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, x_row_stride, rows, V, row_max, row_sumexp,
                                                                 (__nv_bfloat16*)out_bf16, out_row_stride);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_fingerprint(const void* const* ptrs_host, const int64_t* nbytes_host, int n, uint64_t* out, void* stream) {
    TASU_CHECK_ARG(n >= 0 && n <= 8, "at most 8 buffers per call");
    TASU_CHECK_ARG(out != nullptr && (n == 0 || (ptrs_host && nbytes_host)), "null pointer");
    FingerprintArgs a{};
    a.n = n;
    for (int i = 0; i < n; ++i) {
        TASU_CHECK_ARG(nbytes_host[i] >= 0 && (nbytes_host[i] == 0 || ptrs_host[i] != nullptr), "buffer");
        TASU_CHECK_ARG((uintptr_t)ptrs_host[i] % 4 == 0, "4-byte aligned buffers");
        a.ptr[i] = (const uint32_t*)ptrs_host[i];
        a.words[i] = nbytes_host[i] / 4;
        a.stride[i] = a.words[i] <= kFpSamples ? 128 : (a.words[i] - 128) / (kFpSamples / 128 - 1);
    }
    fingerprint_kernel<<<8, kFpThreads, 0, (cudaStream_t)stream>>>(a, (unsigned long long*)out);   // one cluster of 8 CTAs
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_cast_rows_sumsq(const float* src, int64_t rows, int cols, int64_t src_stride, void* dst_bf16,
                                    int64_t dst_stride, float* row_sumsq, void* stream) {
    TASU_CHECK_ARG(rows >= 0 && cols > 0 && src_stride >= cols && dst_stride >= cols, "shape");
    if (rows == 0) return TASU_OK;
    TASU_CHECK_ARG(src && dst_bf16 && row_sumsq, "null pointer");
    const int vec = cols % 8 == 0 && src_stride % 4 == 0 && dst_stride % 8 == 0 && (uintptr_t)src % 16 == 0 &&
                    (uintptr_t)dst_bf16 % 16 == 0;
    int64_t g = (rows + 7) / 8, gmax = (int64_t)tasu::sm_count() * 16;
    if (g > gmax) g = gmax;
    cast_f32_bf16_sumsq_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, src_stride, (__nv_bfloat16*)dst_bf16,
                                                                            dst_stride, dst_stride, row_sumsq, vec);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_kept_frame_index(const int32_t* seg_start, const int32_t* seg_len, const int32_t* seg_frame_off,
                                     const int32_t* row_off, const int32_t* frame_off, int B, int T, int n_prefix,
                                     int64_t max_rows, int64_t max_out, int32_t* frame_row, int32_t* seg_src, void* stream) {
    TASU_CHECK_ARG(B >= 0 && T >= 0 && n_prefix >= 0 && max_rows >= 0 && max_out >= 0, "shape");
    if (B == 0 || max_out == 0) return TASU_OK;
    TASU_CHECK_ARG(seg_start && seg_len && seg_frame_off && row_off && frame_off && frame_row && seg_src, "null pointer");
    kept_frame_index_kernel<<<(unsigned)((max_out + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        seg_start, seg_len, seg_frame_off, row_off, frame_off, B, T, n_prefix, max_rows, max_out, frame_row, seg_src);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
