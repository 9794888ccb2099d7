// Step 2 of the bridge: per-frame greedy statistics over the ~25k-way vocab axis and the
// collapse (run-merge + blank-drop) plan.  Replaces torch.softmax / .max() / .argmax and the
// Python run-length loop of Multitask/model/ps-slm.py:254-301.
//
// frame_stats: one warp per frame, 128-bit streaming loads with a per-row head/tail peel (a
// dense [N, 25055] fp32 row starts 16-byte aligned only every 4th row), U=4 loads in flight
// per lane, online softmax (max / sum-exp) fused with the argmax.  HBM-bound: each element is
// read exactly once.
#include "common.cuh"
#include <limits.h>

namespace tasu {

constexpr float kLog2e = 1.4426950408889634f;

template <int N, bool kLogits>
__device__ __forceinline__ void stats_chunk(const float (&f)[N], int e0, float& best, int& pos, float& s) {
    float cm = f[0];
#pragma unroll
    for (int j = 1; j < N; ++j) cm = fmaxf(cm, f[j]);
    if (cm > best) {                       // rare after the first few chunks
        if (kLogits) s *= exp2f((best - cm) * kLog2e);
        best = cm;
        int j0 = N - 1;
#pragma unroll
        for (int j = N - 2; j >= 0; --j) j0 = (f[j] == cm) ? j : j0;
        pos = e0 + j0;                     // first index of the maximum inside the chunk
    }
    if (kLogits) {
        const float mb = best * kLog2e;
#pragma unroll
        for (int j = 0; j < N; ++j) s += exp2f(fmaf(f[j], kLog2e, -mb));
    }
}

template <typename T, bool kLogits>
__global__ void __launch_bounds__(256)
frame_stats_kernel(const T* __restrict__ x, int B, int T_, int V, int64_t bstride, int64_t rstride,
                   int blank, const int64_t* __restrict__ lens, int32_t* __restrict__ argmax,
                   float* __restrict__ xblank, float* __restrict__ rmax, float* __restrict__ rsum,
                   uint32_t* __restrict__ gmax) {
    constexpr int VEC = Vec16<T>::N;
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (int64_t)B * T_) return;
    const int b = (int)(row / T_), t = (int)(row % T_);
    if (kLogits && lens != nullptr && t >= lens[b]) return;
    const T* p = x + (int64_t)b * bstride + (int64_t)t * rstride;

    const int mis = (int)((reinterpret_cast<uintptr_t>(p) & 15u) / sizeof(T));
    int head = (VEC - mis) % VEC;
    if (head > V) head = V;
    const int nvec = (V - head) / VEC;
    const int tail0 = head + nvec * VEC;

    float best = -INFINITY, s = 0.f;
    int pos = INT_MAX;
    if (lane < head) {
        best = to_f32(p[lane]);
        pos = lane;
        if (kLogits) s = 1.f;
    }
    const uint4* pv = reinterpret_cast<const uint4*>(p + head);
    for (int i = lane; i < nvec; i += 32 * U) {
        uint4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + 32 * u < nvec) q[u] = ld_stream_u4(pv + i + 32 * u);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + 32 * u < nvec) {
                float f[VEC];
                unpack16(q[u], f, T());
                stats_chunk<VEC, kLogits>(f, head + (i + 32 * u) * VEC, best, pos, s);
            }
        }
    }
    if (tail0 + lane < V) {
        float f[1] = {to_f32(p[tail0 + lane])};
        stats_chunk<1, kLogits>(f, tail0 + lane, best, pos, s);
    }
    // cross-lane: maximum with lowest index on ties (torch.argmax rule, ps-slm.py:265)
    float m = best;
    int mp = pos;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, m, o);
        int op = __shfl_xor_sync(0xffffffffu, mp, o);
        if (ob > m || (ob == m && op < mp)) { m = ob; mp = op; }
    }
    if (kLogits) {
        s *= exp2f((best - m) * kLog2e);   // lanes that saw nothing: 0 * exp2(-inf) = 0
        s = warp_sum(s);
    }
    if (lane == 0) {
        argmax[row] = mp;
        xblank[row] = to_f32(p[blank]);
        rmax[row] = m;
        if (kLogits) rsum[row] = s;
        else atomicMax(gmax, float_to_ordered(m));
    }
}

// ---------------------------------------------------------------------------------------------
// Collapse plan: one CTA per utterance.  A frame opens a candidate when it is the first frame,
// a blank, or differs from its left neighbour (ps-slm.py:270-271, :275-279); the opening thread
// walks its run sequentially (same summation order as the reference's .mean(), :286), compares
// the fp32 score with the threshold (:295) and a block scan compacts the kept candidates.
// exclusive scans over the utterances (new_lens → row_off, kept_frames → frame_off) + the plan header; one CTA
__device__ __forceinline__ void collapse_scan_body(const int64_t* __restrict__ new_lens, const int32_t* __restrict__ kept_frames,
                                                   const uint32_t* __restrict__ gmax, int B, int32_t* __restrict__ row_off,
                                                   int32_t* __restrict__ frame_off, int64_t* __restrict__ header,
                                                   int32_t* __restrict__ counts_dev, int* scratch) {
    int carry = 0, mx = 0, frames = 0;
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        const int v = b < B ? (int)__ldcg(new_lens + b) : 0;       // written by other CTAs in the fused variant: bypass L1
        int total;
        const int excl = block_excl_scan_i(v, scratch, &total);
        if (b < B) row_off[b] = carry + excl;
        carry += total;
        mx = max(mx, v);
        if (kept_frames != nullptr) {
            const int kf = b < B ? __ldcg(kept_frames + b) : 0;
            int ftotal;
            const int fexcl = block_excl_scan_i(kf, scratch, &ftotal);
            if (frame_off != nullptr && b < B) frame_off[b] = frames + fexcl;
            frames += ftotal;
        }
    }
    mx = block_max_i(mx, scratch);
    if (frame_off != nullptr && threadIdx.x == 0) frame_off[B] = frames;
    if (threadIdx.x == 0) {
        row_off[B] = carry;
        header[TASU_CH_N_OUT] = carry;
        header[TASU_CH_MAX_LEN] = mx;
        header[TASU_CH_IS_LOGPROB] = (gmax != nullptr && ordered_to_float(*gmax) <= 0.f) ? 1 : 0;
        header[TASU_CH_KEPT_FRAMES] = frames;
        if (counts_dev != nullptr) { counts_dev[0] = carry; counts_dev[1] = mx; counts_dev[2] = frames; counts_dev[3] = 0; }
    }
}

struct CollapseScanArgs { int32_t* ticket; int32_t* row_off; int32_t* frame_off; int64_t* header; int32_t* counts_dev; };

template <bool kLogits>
__global__ void __launch_bounds__(256)
collapse_plan_kernel(const int32_t* __restrict__ ids_all, const float* __restrict__ xb_all,
                     const float* __restrict__ rmax_all, const float* __restrict__ rsum_all,
                     const uint32_t* __restrict__ gmax, const int64_t* __restrict__ lens, int T_,
                     int blank, float thr, int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_len,
                     float* __restrict__ seg_score, int64_t* __restrict__ new_lens, int32_t* __restrict__ kept_frames,
                     int32_t* __restrict__ seg_foff, const CollapseScanArgs sc) {
    __shared__ int scratch[33];
    __shared__ int s_last;
    const int b = blockIdx.x;
    int64_t L64 = lens[b];
    const int L = (int)(L64 < 0 ? 0 : (L64 > T_ ? T_ : L64));
    const int32_t* ids = ids_all + (int64_t)b * T_;
    const float* xb = xb_all + (int64_t)b * T_;
    const float* rmax = kLogits ? rmax_all + (int64_t)b * T_ : nullptr;
    const float* rsum = kLogits ? rsum_all + (int64_t)b * T_ : nullptr;
    const bool is_log = !kLogits && gmax != nullptr && ordered_to_float(*gmax) <= 0.f;   // ps-slm.py:256
    auto pblank = [&](int t) -> float {
        if (kLogits) return expf(xb[t] - rmax[t]) / rsum[t];
        return is_log ? expf(xb[t]) : xb[t];
    };
    int carry = 0, frames = 0;
    for (int t0 = 0; t0 < L; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        int flag = 0, n = 0;
        float score = 0.f;
        if (t < L) {
            const int id = ids[t];
            const bool opens = (t == 0) || (id == blank) || (ids[t - 1] != id);
            if (opens) {
                float sum = pblank(t);
                n = 1;
                if (id != blank)
                    while (t + n < L && ids[t + n] == id) { sum += pblank(t + n); ++n; }
                score = (n == 1) ? sum : sum / (float)n;
                flag = score < thr;
            }
        }
        int total, ftotal;
        const int excl = block_excl_scan_i(flag, scratch, &total);
        const int fexcl = block_excl_scan_i(flag ? n : 0, scratch, &ftotal);
        if (flag) {
            const int64_t j = (int64_t)b * T_ + carry + excl;
            seg_start[j] = t;
            seg_len[j] = n;
            if (seg_score) seg_score[j] = score;
            if (seg_foff) seg_foff[j] = frames + fexcl;     // kept-frame offset inside the utterance
        }
        carry += total;
        frames += ftotal;
    }
    if (kept_frames != nullptr && threadIdx.x == 0) kept_frames[b] = frames;
    if (threadIdx.x == 0) new_lens[b] = carry;
    if (sc.ticket == nullptr) return;
    // fused scan (tasu_collapse_plan_scan): the last CTA to finish sees every utterance's counts and writes the offsets
    // and the header; the ticket returns to zero, so the same word serves the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int t = atomicAdd(sc.ticket, 1);
        s_last = (t == (int)gridDim.x - 1);
        if (s_last) { *sc.ticket = 0; __threadfence(); }
    }
    __syncthreads();
    if (!s_last) return;
    collapse_scan_body(new_lens, kept_frames, kLogits ? nullptr : gmax, (int)gridDim.x, sc.row_off, sc.frame_off, sc.header,
                       sc.counts_dev, scratch);
}

__global__ void __launch_bounds__(1024)
collapse_scan_kernel(const int64_t* __restrict__ new_lens, const int32_t* __restrict__ kept_frames,
                     const uint32_t* __restrict__ gmax, int B, int32_t* __restrict__ row_off,
                     int32_t* __restrict__ frame_off, int64_t* __restrict__ header, int32_t* __restrict__ counts_dev) {
    __shared__ int scratch[33];
    collapse_scan_body(new_lens, kept_frames, gmax, B, row_off, frame_off, header, counts_dev, scratch);
}

}  // namespace tasu

using namespace tasu;

extern "C" int tasu_frame_stats(const void* x, int dtype, int input_kind, int B, int T, int V,
                                int64_t batch_stride, int64_t row_stride, int blank_id,
                                const int64_t* lens, int32_t* argmax, float* x_blank, float* row_max,
                                float* row_sumexp, uint32_t* global_max_enc, void* stream) {
    TASU_CHECK_ARG(B >= 0 && T >= 0 && V > 0, "B,T >= 0 and V > 0");
    TASU_CHECK_ARG(blank_id >= 0 && blank_id < V, "blank_id out of range");
    TASU_CHECK_ARG(dtype == TASU_F32 || dtype == TASU_BF16, "dtype");
    TASU_CHECK_ARG(input_kind == TASU_INPUT_PROBS || input_kind == TASU_INPUT_LOGITS, "input_kind");
    TASU_CHECK_ARG(argmax && x_blank && row_max, "null output");
    TASU_CHECK_ARG(input_kind == TASU_INPUT_PROBS ? global_max_enc != nullptr : row_sumexp != nullptr,
                   "global_max_enc (probs) / row_sumexp (logits) required");
    cudaStream_t st = (cudaStream_t)stream;
    if (global_max_enc) TASU_CHECK_CUDA(cudaMemsetAsync(global_max_enc, 0, sizeof(uint32_t), st));
    const int64_t rows = (int64_t)B * T;
    if (rows == 0) return TASU_OK;
    TASU_CHECK_ARG(x != nullptr, "null input");
    const int wpb = 8;
    const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
#define LAUNCH(TT, LG)                                                                                   \
    frame_stats_kernel<TT, LG><<<grid, wpb * 32, 0, st>>>((const TT*)x, B, T, V, batch_stride, row_stride, \
                                                          blank_id, lens, argmax, x_blank, row_max,        \
                                                          row_sumexp, global_max_enc)
    if (dtype == TASU_F32) {
        if (input_kind == TASU_INPUT_LOGITS) LAUNCH(float, true); else LAUNCH(float, false);
    } else {
        if (input_kind == TASU_INPUT_LOGITS) LAUNCH(__nv_bfloat16, true); else LAUNCH(__nv_bfloat16, false);
    }
#undef LAUNCH
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

static int launch_collapse_plan(const int32_t* argmax, const float* x_blank, const float* row_max, const float* row_sumexp,
                                const uint32_t* global_max_enc, int input_kind, const int64_t* lens, int B, int T, int blank_id,
                                float threshold, int32_t* seg_start, int32_t* seg_len, float* seg_score, int64_t* new_lens,
                                int32_t* kept_frames, int32_t* seg_frame_off, const CollapseScanArgs& sc, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (input_kind == TASU_INPUT_LOGITS)
        collapse_plan_kernel<true><<<B, 256, 0, st>>>(argmax, x_blank, row_max, row_sumexp, global_max_enc, lens, T, blank_id,
                                                      threshold, seg_start, seg_len, seg_score, new_lens, kept_frames, seg_frame_off, sc);
    else
        collapse_plan_kernel<false><<<B, 256, 0, st>>>(argmax, x_blank, row_max, row_sumexp, global_max_enc, lens, T, blank_id,
                                                       threshold, seg_start, seg_len, seg_score, new_lens, kept_frames, seg_frame_off, sc);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_collapse_plan(const int32_t* argmax, const float* x_blank, const float* row_max,
                                  const float* row_sumexp, const uint32_t* global_max_enc, int input_kind,
                                  const int64_t* lens, int B, int T, int blank_id, float threshold,
                                  int32_t* seg_start, int32_t* seg_len, float* seg_score,
                                  int64_t* new_lens, int32_t* kept_frames, int32_t* seg_frame_off, void* stream) {
    TASU_CHECK_ARG(B >= 0 && T >= 0, "B,T >= 0");
    TASU_CHECK_ARG(input_kind == TASU_INPUT_PROBS || input_kind == TASU_INPUT_LOGITS, "input_kind");
    if (B == 0) return TASU_OK;
    TASU_CHECK_ARG(lens && seg_start && seg_len && new_lens, "null pointer");
    TASU_CHECK_ARG(T == 0 || (argmax && x_blank), "null stats");
    TASU_CHECK_ARG(input_kind != TASU_INPUT_LOGITS || (row_max && row_sumexp), "logits need row_max/row_sumexp");
    return launch_collapse_plan(argmax, x_blank, row_max, row_sumexp, global_max_enc, input_kind, lens, B, T, blank_id, threshold,
                                seg_start, seg_len, seg_score, new_lens, kept_frames, seg_frame_off, CollapseScanArgs{}, stream);
}

extern "C" int tasu_collapse_plan_scan(const int32_t* argmax, const float* x_blank, const float* row_max,
                                       const float* row_sumexp, const uint32_t* global_max_enc, int input_kind,
                                       const int64_t* lens, int B, int T, int blank_id, float threshold,
                                       int32_t* seg_start, int32_t* seg_len, float* seg_score, int64_t* new_lens,
                                       int32_t* kept_frames, int32_t* seg_frame_off, int32_t* row_off, int32_t* frame_off,
                                       int64_t* header, int32_t* counts_dev, int32_t* ticket, void* stream) {
    TASU_CHECK_ARG(B >= 0 && T >= 0, "B,T >= 0");
    TASU_CHECK_ARG(input_kind == TASU_INPUT_PROBS || input_kind == TASU_INPUT_LOGITS, "input_kind");
    TASU_CHECK_ARG(row_off && header && ticket, "null scan outputs / ticket");
    if (B == 0)
        return tasu_collapse_scan(new_lens, kept_frames, input_kind == TASU_INPUT_PROBS ? global_max_enc : nullptr, 0, row_off,
                                  frame_off, header, counts_dev, stream);
    TASU_CHECK_ARG(lens && seg_start && seg_len && new_lens, "null pointer");
    TASU_CHECK_ARG(T == 0 || (argmax && x_blank), "null stats");
    TASU_CHECK_ARG(input_kind != TASU_INPUT_LOGITS || (row_max && row_sumexp), "logits need row_max/row_sumexp");
    CollapseScanArgs sc{ticket, row_off, frame_off, header, counts_dev};
    return launch_collapse_plan(argmax, x_blank, row_max, row_sumexp, global_max_enc, input_kind, lens, B, T, blank_id, threshold,
                                seg_start, seg_len, seg_score, new_lens, kept_frames, seg_frame_off, sc, stream);
}

extern "C" int tasu_collapse_scan(const int64_t* new_lens, const int32_t* kept_frames,
                                  const uint32_t* global_max_enc, int B, int32_t* row_off, int32_t* frame_off,
                                  int64_t* header, int32_t* counts_dev, void* stream) {
    TASU_CHECK_ARG(B >= 0 && row_off && header, "B >= 0, non-null outputs");
    TASU_CHECK_ARG(B == 0 || new_lens, "null new_lens");
    collapse_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(new_lens, kept_frames, global_max_enc, B, row_off, frame_off, header, counts_dev);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
