// Exact greedy decisions for the fused CTC head (tasu_ctc_head_stats computes its logits from bf16 operands).
//
// The integers of PSD — argmax per frame (Multitask/model/ps-slm.py:265) and the strict fp32 `score < 0.9` test
// (:295-297) — must equal the fp32 reference.  A bf16 logit differs from the fp32 one by at most
//     delta_f = ||x_f||_2 · max_v ||w_v||_2 · 2^-8          (both operands rounded to 8 significant bits, Cauchy-Schwarz)
// so a decision of frame f can only differ when its margin is below 2·delta_f:
//   * argmax: the gap between the two largest logits.  A lower bound of the gap comes for free from the statistics the
//     head already emits: p1 = 1/Σexp, and the runner-up probability is at most min(1 - p1, sqrt(Σp² - p1²));
//   * keep/drop: logit(p_blank) moves by at most 2·delta_f, so a blank frame is decided unless
//     |logit(p_blank) - logit(thr)| < 2·delta_f, and a non-blank run stays below the threshold whenever every one of its
//     frames does (its score is the mean of the frames' blank probabilities, :286).
// flag_ambiguous_kernel lists the frames inside those margins (an uncapped list: one slot per frame exists), and
// ctc_refine_kernel recomputes exactly those frames with fp32 FMAs on the ORIGINAL fp32 weights and fp32 (or bf16-given)
// encoder rows, looping over the whole list in-kernel — cost proportional to the number of listed frames, no host
// involvement, no capacity to overflow.  The refined (argmax, blank logit, max, Σexp) replace the head's values for the
// collapse plan only; the softmax normalisers of pass 2 stay the bf16-consistent ones.
#include "common.cuh"
#include <limits.h>
#include <algorithm>

namespace tasu {

template <typename T>
__global__ void __launch_bounds__(256)
row_norm_max_kernel(const T* __restrict__ w, int rows, int cols, int64_t ld, uint32_t* __restrict__ out_enc) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const T* p = w + r * ld;
    float ss = 0.f;
    for (int k = lane; k < cols; k += 32) { const float v = to_f32(p[k]); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    if (lane == 0) atomicMax(out_enc, float_to_ordered(sqrtf(ss)));
}

struct FlagArgs {
    const int32_t* argmax; const float* x_blank; const float* row_max; const float* row_sumexp; const float* row_sumexp2;
    const int64_t* lens; const uint32_t* w_norm_max_enc; float err_scale; int B, T, n_prefix, blank; float logit_thr;
    float* dec_max; float* dec_sum; int32_t* frame_idx; int32_t* raw_row; int32_t* count;
};

// decision of one valid frame f (raw row r) whose squared input norm is ss; called by one thread
__device__ __forceinline__ void flag_frame(const FlagArgs& a, int64_t f, int64_t r, float ss) {
    const float m = a.row_max[f], s = a.row_sumexp[f];
    const float margin = 2.f * a.err_scale * sqrtf(ss) * ordered_to_float(*a.w_norm_max_enc);
    // lower bound of the top-2 logit gap: log(p1 / p2_max)
    const float p1 = 1.f / s;
    float p2 = (s - 1.f) / s;                                      // everything that is not the maximum
    if (a.row_sumexp2 != nullptr) {
        const float q = fmaxf(a.row_sumexp2[f] / (s * s) - p1 * p1, 0.f) + 4e-6f;  // Σ_{v != argmax} p_v² (+ rounding slack)
        p2 = fminf(p2, sqrtf(q));
    }
    const float gap = p2 > 0.f ? logf(p1) - logf(p2) : INFINITY;
    // logit of the blank probability
    const float xb = a.x_blank[f];
    const float rest = s - expf(xb - m);
    const float lb = rest > 0.f ? (xb - m) - logf(rest) : INFINITY;
    bool amb = !(gap >= margin);                                   // NaN statistics are listed, never trusted
    if (a.argmax[f] == a.blank) amb = amb || fabsf(lb - a.logit_thr) < margin;
    else amb = amb || (lb + margin >= a.logit_thr);
    if (!amb) return;
    const int k = atomicAdd(a.count, 1);
    a.frame_idx[k] = (int32_t)f;
    a.raw_row[k] = (int32_t)r;
}

// one warp per frame: the warp sums the squares of the frame's encoder row
template <typename TX>
__global__ void __launch_bounds__(256)
flag_ambiguous_kernel(const FlagArgs a, const TX* __restrict__ x, int64_t ldx, int K) {
    const int lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= (int64_t)a.B * a.T) return;
    const int b = (int)(f / a.T), t = (int)(f % a.T);
    if (lane == 0) { a.dec_max[f] = a.row_max[f]; a.dec_sum[f] = a.row_sumexp[f]; }   // the plan's normalisers; refined frames are overwritten
    if (t >= a.lens[b]) return;
    const int64_t r = (int64_t)b * (a.T + a.n_prefix) + a.n_prefix + t;
    const TX* px = x + r * ldx;
    float ss = 0.f;
    for (int k = lane; k < K; k += 32) { const float v = to_f32(px[k]); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    if (lane == 0) flag_frame(a, f, r, ss);
}

// one thread per frame: the squared row norms were produced by the cast of the encoder output (tasu_cast_rows_sumsq)
__global__ void __launch_bounds__(256)
flag_ambiguous_sumsq_kernel(const FlagArgs a, const float* __restrict__ x_sumsq) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= (int64_t)a.B * a.T) return;
    const int b = (int)(f / a.T), t = (int)(f % a.T);
    a.dec_max[f] = a.row_max[f];
    a.dec_sum[f] = a.row_sumexp[f];
    if (t >= a.lens[b]) return;
    const int64_t r = (int64_t)b * (a.T + a.n_prefix) + a.n_prefix + t;
    flag_frame(a, f, r, x_sumsq[r]);
}

// ---------------------------------------------------------------------------------------------------------------------
// fp32 recomputation of the listed frames.  Work item = (chunk of 32 listed frames, vocabulary split); a CTA computes
// 32 frames x 256 classes per tile with a 4 x 8 register tile per thread (warp w owns frames 4w..4w+3 of the chunk, lane l
// the classes n0 + l + 32j), operands staged through shared memory in K chunks of 32 with the next chunk prefetched into
// registers, every dot product accumulated in ascending k with fp32 FMAs.
constexpr int kRfRows = 32, kRfCols = 256, kRfKc = 32, kRfPitch = kRfKc + 4, kRfThreads = 256;

struct RefineSched { int n, n_chunks, n_tiles, splits, tiles_per, items; };

__device__ __forceinline__ RefineSched refine_schedule(int count, int max_frames, int V, int grid) {
    RefineSched sc;
    sc.n = min(max(count, 0), max_frames);
    sc.n_chunks = (sc.n + kRfRows - 1) / kRfRows;
    sc.n_tiles = (V + kRfCols - 1) / kRfCols;
    int s = sc.n_chunks >= grid ? 1 : grid / max(sc.n_chunks, 1);
    s = max(1, min(s, sc.n_tiles));
    sc.tiles_per = (sc.n_tiles + s - 1) / s;
    sc.splits = (sc.n_tiles + sc.tiles_per - 1) / sc.tiles_per;
    sc.items = sc.n_chunks * sc.splits;
    return sc;
}

struct RefineParams {
    const void* x; int64_t ldx; const float* w; int64_t ldw; const float* bias; int V, K, blank;
    const int32_t* raw_row; const int32_t* count; int max_frames;
    float* part_max; float* part_sum; float* part_xb; int32_t* part_arg;
};

template <typename TX>
__device__ __forceinline__ float4 load_x4(const TX* p, int k, int K, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && k + 3 < K) {
        if constexpr (sizeof(TX) == 4) {
            v = *reinterpret_cast<const float4*>(p + k);
        } else {
            const uint2 q = *reinterpret_cast<const uint2*>(p + k);
            v.x = __uint_as_float(q.x << 16); v.y = __uint_as_float(q.x & 0xffff0000u);
            v.z = __uint_as_float(q.y << 16); v.w = __uint_as_float(q.y & 0xffff0000u);
        }
    } else {
        if (k < K) v.x = to_f32(p[k]);
        if (k + 1 < K) v.y = to_f32(p[k + 1]);
        if (k + 2 < K) v.z = to_f32(p[k + 2]);
        if (k + 3 < K) v.w = to_f32(p[k + 3]);
    }
    return v;
}

template <typename TX>
__global__ void __launch_bounds__(kRfThreads, 2)
ctc_refine_kernel(const RefineParams p, int vec_x, int vec_w) {
    __shared__ __align__(16) float xs[kRfRows * kRfPitch];
    __shared__ __align__(16) float ws[kRfCols * kRfPitch];
    const RefineSched sc = refine_schedule(*p.count, p.max_frames, p.V, (int)gridDim.x);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TX* x = reinterpret_cast<const TX*>(p.x);
    const int xr = tid >> 3, xk = (tid & 7) * 4;                   // this thread's slot of the x chunk (row, k quad)
    for (int item = blockIdx.x; item < sc.items; item += gridDim.x) {
        const int chunk = item / sc.splits, split = item % sc.splits;
        const int row0 = chunk * kRfRows;
        const int t0 = split * sc.tiles_per, t1 = min(sc.n_tiles, t0 + sc.tiles_per);
        const int my_slot = row0 + xr;
        const TX* my_x = my_slot < sc.n ? x + (int64_t)p.raw_row[my_slot] * p.ldx : nullptr;
        float rm[4], rs[4], rxb[4];
        int ra[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { rm[r] = -INFINITY; rs[r] = 0.f; rxb[r] = -INFINITY; ra[r] = INT_MAX; }
        for (int tile = t0; tile < t1; ++tile) {
            const int n0 = tile * kRfCols;
            float acc[4][8];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
            float4 px, pw[8];
            auto fetch = [&](int k0) {
                px = my_x ? load_x4<TX>(my_x, k0 + xk, p.K, vec_x != 0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = tid + kRfThreads * i, col = n0 + (idx >> 3), kq = (idx & 7) * 4;
                    pw[i] = col < p.V ? load_x4<float>(p.w + (int64_t)col * p.ldw, k0 + kq, p.K, vec_w != 0)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            fetch(0);
            for (int k0 = 0; k0 < p.K; k0 += kRfKc) {
                __syncthreads();                                   // the previous chunk has been consumed
                *reinterpret_cast<float4*>(&xs[xr * kRfPitch + xk]) = px;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = tid + kRfThreads * i;
                    *reinterpret_cast<float4*>(&ws[(idx >> 3) * kRfPitch + (idx & 7) * 4]) = pw[i];
                }
                __syncthreads();
                if (k0 + kRfKc < p.K) fetch(k0 + kRfKc);           // in flight while this chunk is multiplied
#pragma unroll
                for (int kk = 0; kk < kRfKc; kk += 4) {
                    float4 xv[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) xv[r] = *reinterpret_cast<const float4*>(&xs[(warp * 4 + r) * kRfPitch + kk]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 wv = *reinterpret_cast<const float4*>(&ws[(lane + 32 * j) * kRfPitch + kk]);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {              // ascending k, one rounding per term
                            acc[r][j] = fmaf(xv[r].x, wv.x, acc[r][j]);
                            acc[r][j] = fmaf(xv[r].y, wv.y, acc[r][j]);
                            acc[r][j] = fmaf(xv[r].z, wv.z, acc[r][j]);
                            acc[r][j] = fmaf(xv[r].w, wv.w, acc[r][j]);
                        }
                    }
                }
            }
            // online softmax statistics of this tile (classes ascend with j, tiles ascend: strict > keeps the first maximum)
            float bj[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = n0 + lane + 32 * j;
                bj[j] = col < p.V ? (p.bias ? __ldg(p.bias + col) : 0.f) : -INFINITY;
            }
            const int bl = p.blank - n0;                           // blank column inside this tile?
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float cm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc[r][j] += bj[j]; cm = fmaxf(cm, acc[r][j]); }
                if (cm > rm[r]) {
                    rs[r] *= expf(rm[r] - cm);                     // exp(-inf) = 0 on the first tile
                    rm[r] = cm;
                    int j0 = 7;
#pragma unroll
                    for (int j = 6; j >= 0; --j) j0 = (acc[r][j] == cm) ? j : j0;
                    ra[r] = n0 + lane + 32 * j0;
                }
                if (rm[r] > -INFINITY) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) rs[r] += expf(acc[r][j] - rm[r]);
                }
                if (bl >= 0 && bl < kRfCols && (bl & 31) == lane) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) rxb[r] = (j == (bl >> 5)) ? acc[r][j] : rxb[r];
                }
            }
        }
        // merge the 32 lanes of every frame: maximum with the lowest class index on ties
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float m = rm[r], s = rs[r];
            int a = ra[r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
                const int oa = __shfl_xor_sync(0xffffffffu, a, o);
                const float nm = fmaxf(m, om);
                s = (nm > -INFINITY) ? s * expf(m - nm) + os * expf(om - nm) : 0.f;
                a = (om > m || (om == m && oa < a)) ? oa : a;
                m = nm;
            }
            const float xb = warp_max(rxb[r]);
            const int slot = row0 + warp * 4 + r;
            if (lane == 0 && slot < sc.n) {
                const int64_t o = (int64_t)split * (sc.n_chunks * kRfRows) + slot;
                p.part_max[o] = m; p.part_sum[o] = s; p.part_arg[o] = a; p.part_xb[o] = xb;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
ctc_refine_combine_kernel(const RefineParams p, int refine_grid, const int32_t* __restrict__ frame_idx,
                          int32_t* __restrict__ argmax, float* __restrict__ x_blank, float* __restrict__ dec_max,
                          float* __restrict__ dec_sum) {
    const RefineSched sc = refine_schedule(*p.count, p.max_frames, p.V, refine_grid);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= sc.n) return;
    const int64_t stride = (int64_t)sc.n_chunks * kRfRows;
    float m = -INFINITY;
    int a = INT_MAX;
    for (int s = 0; s < sc.splits; ++s) {
        const float pm = p.part_max[s * stride + k];
        if (pm > m) { m = pm; a = p.part_arg[s * stride + k]; }    // ties keep the lower split = lower class index
    }
    float sum = 0.f;
    for (int s = 0; s < sc.splits; ++s) sum += p.part_sum[s * stride + k] * expf(p.part_max[s * stride + k] - m);
    const int sb = (p.blank / kRfCols) / sc.tiles_per;             // the split that saw the blank column
    const int f = frame_idx[k];
    argmax[f] = a;
    x_blank[f] = p.part_xb[sb * stride + k];
    dec_max[f] = m;
    dec_sum[f] = sum;
}

static int refine_grid() { return 2 * sm_count(); }

}  // namespace tasu

using namespace tasu;

extern "C" int tasu_row_norm_max(const void* w, int dtype, int rows, int cols, int64_t ld, uint32_t* out_enc, void* stream) {
    TASU_CHECK_ARG(rows >= 0 && cols >= 0 && ld >= cols, "shape");
    TASU_CHECK_ARG(dtype == TASU_F32 || dtype == TASU_BF16, "dtype");
    TASU_CHECK_ARG(out_enc != nullptr, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    TASU_CHECK_CUDA(cudaMemsetAsync(out_enc, 0, sizeof(uint32_t), st));
    if (rows == 0) return TASU_OK;
    TASU_CHECK_ARG(w != nullptr, "null input");
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (dtype == TASU_F32) row_norm_max_kernel<float><<<grid, 256, 0, st>>>((const float*)w, rows, cols, ld, out_enc);
    else row_norm_max_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)w, rows, cols, ld, out_enc);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_flag_ambiguous_frames(const int32_t* argmax, const float* x_blank, const float* row_max,
                                          const float* row_sumexp, const float* row_sumexp2, const int64_t* lens,
                                          const void* x, int x_dtype, int64_t ldx, int K, const float* x_sumsq,
                                          const uint32_t* w_norm_max_enc, float err_scale, int B, int T, int n_prefix,
                                          int blank_id, float threshold, float* dec_max, float* dec_sum, int32_t* frame_idx,
                                          int32_t* raw_row, int32_t* count, void* stream) {
    TASU_CHECK_ARG(B >= 0 && T >= 0 && n_prefix >= 0 && K > 0 && ldx >= K, "shape");
    TASU_CHECK_ARG(x_dtype == TASU_F32 || x_dtype == TASU_BF16, "x_dtype");
    TASU_CHECK_ARG(err_scale >= 0.f, "err_scale");
    TASU_CHECK_ARG(count != nullptr, "null count");
    TASU_CHECK_ARG((int64_t)B * (T + n_prefix) < (1LL << 31), "too many frames for one call");
    cudaStream_t st = (cudaStream_t)stream;
    TASU_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    const int64_t n = (int64_t)B * T;
    if (n == 0) return TASU_OK;
    TASU_CHECK_ARG(argmax && x_blank && row_max && row_sumexp && lens && (x || x_sumsq) && w_norm_max_enc && dec_max && dec_sum &&
                   frame_idx && raw_row, "null pointer");
    FlagArgs a{};
    a.argmax = argmax; a.x_blank = x_blank; a.row_max = row_max; a.row_sumexp = row_sumexp; a.row_sumexp2 = row_sumexp2;
    a.lens = lens; a.w_norm_max_enc = w_norm_max_enc; a.err_scale = err_scale; a.B = B; a.T = T; a.n_prefix = n_prefix;
    a.blank = blank_id;
    a.logit_thr = threshold >= 1.f ? INFINITY : threshold <= 0.f ? -INFINITY : logf(threshold) - log1pf(-threshold);
    a.dec_max = dec_max; a.dec_sum = dec_sum; a.frame_idx = frame_idx; a.raw_row = raw_row; a.count = count;
    if (x_sumsq != nullptr)
        flag_ambiguous_sumsq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, x_sumsq);
    else if (x_dtype == TASU_F32)
        flag_ambiguous_kernel<float><<<(unsigned)((n + 7) / 8), 256, 0, st>>>(a, (const float*)x, ldx, K);
    else
        flag_ambiguous_kernel<__nv_bfloat16><<<(unsigned)((n + 7) / 8), 256, 0, st>>>(a, (const __nv_bfloat16*)x, ldx, K);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int64_t tasu_ctc_head_refine_workspace(int64_t max_frames) {
    const int64_t entries = std::max<int64_t>((int64_t)refine_grid() * kRfRows, max_frames + kRfRows);
    return entries * 16 + 256;
}

extern "C" int tasu_ctc_head_refine(const void* x, int x_dtype, int64_t ldx, const float* w_f32, int64_t ldw,
                                    const float* bias, int V, int K, int blank_id, const int32_t* frame_idx,
                                    const int32_t* raw_row, const int32_t* count, int64_t max_frames, int32_t* argmax,
                                    float* x_blank, float* dec_max, float* dec_sum, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
    TASU_CHECK_ARG(V > 0 && K > 0 && ldx >= K && ldw >= K && max_frames >= 0, "shape");
    TASU_CHECK_ARG(max_frames < (1LL << 31), "too many frames for one call");
    TASU_CHECK_ARG(blank_id >= 0 && blank_id < V, "blank_id out of range");
    TASU_CHECK_ARG(x_dtype == TASU_F32 || x_dtype == TASU_BF16, "x_dtype");
    if (max_frames == 0) return TASU_OK;
    TASU_CHECK_ARG(x && w_f32 && frame_idx && raw_row && count && argmax && x_blank && dec_max && dec_sum && workspace,
                   "null pointer");
    TASU_CHECK_ARG(workspace_bytes >= tasu_ctc_head_refine_workspace(max_frames), "workspace too small");
    TASU_CHECK_ARG((uintptr_t)workspace % 16 == 0, "workspace alignment");
    const int grid = refine_grid();
    const int64_t entries = std::max<int64_t>((int64_t)grid * kRfRows, max_frames + kRfRows);
    RefineParams p{};
    p.x = x; p.ldx = ldx; p.w = w_f32; p.ldw = ldw; p.bias = bias; p.V = V; p.K = K; p.blank = blank_id;
    p.raw_row = raw_row; p.count = count; p.max_frames = (int)max_frames;
    float* ws = reinterpret_cast<float*>(workspace);
    p.part_max = ws; p.part_sum = ws + entries; p.part_xb = ws + 2 * entries;
    p.part_arg = reinterpret_cast<int32_t*>(ws + 3 * entries);
    const size_t xe = x_dtype == TASU_F32 ? 4 : 2;
    const int vec_x = ((uintptr_t)x % 16 == 0) && ((ldx * xe) % 16 == 0);
    const int vec_w = ((uintptr_t)w_f32 % 16 == 0) && ((ldw * 4) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == TASU_F32) ctc_refine_kernel<float><<<grid, kRfThreads, 0, st>>>(p, vec_x, vec_w);
    else ctc_refine_kernel<__nv_bfloat16><<<grid, kRfThreads, 0, st>>>(p, vec_x, vec_w);
    TASU_CHECK_LAUNCH();
    // one thread per list slot; the live count is read on the device
    ctc_refine_combine_kernel<<<(unsigned)((max_frames + 255) / 256), 256, 0, st>>>(p, grid, frame_idx, argmax, x_blank, dec_max, dec_sum);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
