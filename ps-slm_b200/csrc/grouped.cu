// Step 2 (ps-slm.py:275-297), grouped layout of the kept frames: the mean over the frames of a multi-frame run is taken
// INSIDE the epilogue of the kept-frame softmax GEMM (gemm_sm100.cu, kGrouped), so the per-frame probabilities of such
// runs are never written to HBM and tasu_pool_tail has nothing left to do for them.
//
// The frames of a run are laid out on ADJACENT rows of the GEMM's A operand, runs of one size class filling whole
// 128-row tiles:
//     class S  (1 frame)            one row per candidate                       A rows [0, NS)
//     class G2 (2 frames)           2 adjacent rows per candidate               A rows [A2, A2 + 2*N2)
//     class G4 (3-4 frames)         4 adjacent rows (a 3-frame run gets one zero row of weight 0)   [A4, A4 + 4*N4)
//     class X  (> 4 frames, rare)   first frame among the S rows, extra frames in [AX, AX + NXE): written per frame and
//                                   averaged by tasu_pool_tail exactly as in the plain layout
// (A2, A4, AX are multiples of 128; the rows between the regions are zero rows of weight 0).  Accumulator row i of a tile
// lives in TMEM lane i, i.e. in thread i of the epilogue: the 2 / 4 frames of a run sit in adjacent lanes of one warp
// and their probabilities meet through 16 / 24 warp shuffles per 32-column slab (recursive halving: every lane ends up
// with the sum of 16 / 8 columns).  A G2 tile therefore stores 64 pooled rows, a G4 tile 32, each with ONE dense TMA
// store per 64-column chunk.  Pooled rows come out in class order; `perm` maps packed candidate r → pooled row, and the
// projector's output is brought back to packed order by one row gather (25 MB) at the end.
#include "common.cuh"

namespace tasu {

struct GroupScanArgs { int32_t* ticket; int32_t* cnt; int32_t* base; int32_t* lay; };

__device__ __forceinline__ int run_class(int n) { return n <= 1 ? 0 : (n == 2 ? 1 : (n <= 4 ? 2 : 3)); }
__device__ __forceinline__ int round_up_128(int v) { return (v + 127) & ~127; }

// One CTA per utterance: the index of every kept candidate inside its size class (exclusive scan in candidate order →
// the layout is deterministic); the last CTA to finish scans the per-utterance counts and writes the layout words.
// cnt / base: [4][B] = {S and X candidates, G2, G4, extra frames of X}.
__global__ void __launch_bounds__(256)
group_plan_kernel(const int32_t* __restrict__ seg_len, const int64_t* __restrict__ new_lens, int T, int32_t* __restrict__ slot,
                  int32_t* __restrict__ xoff, const GroupScanArgs sc) {
    __shared__ int scratch[33];
    __shared__ int s_last;
    const int b = blockIdx.x, B = gridDim.x;
    const int M = (int)min((int64_t)T, max((int64_t)0, new_lens[b]));
    int c0 = 0, c1 = 0, c2 = 0, cx = 0;
    for (int j0 = 0; j0 < M; j0 += blockDim.x) {
        const int j = j0 + threadIdx.x;
        const int n = j < M ? seg_len[(int64_t)b * T + j] : 0;
        const int cls = j < M ? run_class(n) : -1;
        int t0, t1, t2, tx;
        const int e0 = block_excl_scan_i(cls == 0 || cls == 3, scratch, &t0);
        const int e1 = block_excl_scan_i(cls == 1, scratch, &t1);
        const int e2 = block_excl_scan_i(cls == 2, scratch, &t2);
        const int ex = block_excl_scan_i(cls == 3 ? n - 1 : 0, scratch, &tx);
        if (j < M) {
            slot[(int64_t)b * T + j] = cls == 1 ? c1 + e1 : (cls == 2 ? c2 + e2 : c0 + e0);
            xoff[(int64_t)b * T + j] = cx + ex;
        }
        c0 += t0; c1 += t1; c2 += t2; cx += tx;
    }
    if (threadIdx.x == 0) { sc.cnt[b] = c0; sc.cnt[B + b] = c1; sc.cnt[2 * B + b] = c2; sc.cnt[3 * B + b] = cx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int t = atomicAdd(sc.ticket, 1);
        s_last = (t == B - 1);
        if (s_last) { *sc.ticket = 0; __threadfence(); }
    }
    __syncthreads();
    if (!s_last) return;
    int tot[4];
    for (int c = 0; c < 4; ++c) {
        int carry = 0;
        for (int b0 = 0; b0 < B; b0 += blockDim.x) {
            const int bb = b0 + threadIdx.x;
            const int v = bb < B ? __ldcg(sc.cnt + c * B + bb) : 0;
            int total;
            const int excl = block_excl_scan_i(v, scratch, &total);
            if (bb < B) sc.base[c * B + bb] = carry + excl;
            carry += total;
        }
        tot[c] = carry;
    }
    if (threadIdx.x == 0) {
        const int ns = tot[0], n2 = tot[1], n4 = tot[2], nxe = tot[3];
        const int a2 = round_up_128(ns), a4 = round_up_128(a2 + 2 * n2), ax = round_up_128(a4 + 4 * n4);
        const int o4 = a2 + ((n2 + 63) & ~63), ox = o4 + ((n4 + 31) & ~31);
        int32_t* lay = sc.lay;
        lay[TASU_GL_A2] = a2; lay[TASU_GL_A4] = a4; lay[TASU_GL_AX] = ax; lay[TASU_GL_A_ROWS] = ax + nxe;
        lay[TASU_GL_O4] = o4; lay[TASU_GL_OX] = ox; lay[TASU_GL_N2] = n2; lay[TASU_GL_N4] = n4;
        lay[TASU_GL_NS] = ns; lay[TASU_GL_NXE] = nxe;
        for (int i = TASU_GL_NXE + 1; i < TASU_GL_WORDS; ++i) lay[i] = 0;
    }
}

__device__ __forceinline__ void copy_row_bf16(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int K,
                                              bool vec, int lane) {
    if (vec) {
        for (int c = lane; c < K / 8; c += 32) reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(src)[c];
    } else {
        for (int c = lane; c < K; c += 32) dst[c] = src[c];
    }
}
__device__ __forceinline__ void zero_row_bf16(__nv_bfloat16* __restrict__ dst, int K, bool vec, int lane) {
    if (vec) {
        for (int c = lane; c < K / 8; c += 32) reinterpret_cast<uint4*>(dst)[c] = make_uint4(0u, 0u, 0u, 0u);
    } else {
        for (int c = lane; c < K; c += 32) dst[c] = __float2bfloat16_rn(0.f);
    }
}

// One warp per packed candidate (then per filler row): copy its frames' encoder rows to their place in the grouped
// layout, with the softmax scalars of every A row (g_inv = 1 / (sum exp * frames averaged in the epilogue); 0 = a row
// of weight 0).  max_a / max_o = rows of the A matrix / of the pooled matrix and its per-row arrays; max_out = rows of
// `perm`, max_proj = pooled rows the projector's buffers hold: capacities bound every write (an overflowing batch is redone by the host with larger buffers).
__global__ void __launch_bounds__(256)
gather_grouped_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int B, int T, int n_prefix, int K, int V,
                      const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_len,
                      const int32_t* __restrict__ row_off, const int32_t* __restrict__ slot, const int32_t* __restrict__ xoff,
                      const int32_t* __restrict__ base, const int32_t* __restrict__ lay, const float* __restrict__ row_max,
                      const float* __restrict__ row_sumexp, const float* __restrict__ row_sumexp2, int64_t max_a,
                      int64_t max_o, int64_t max_out, int64_t max_proj, __nv_bfloat16* __restrict__ xg, int64_t ldg, float* __restrict__ g_max,
                      float* __restrict__ g_inv, int32_t* __restrict__ perm, int32_t* __restrict__ pk_len,
                      int32_t* __restrict__ tail_src, int32_t* __restrict__ multi_rows, int32_t* __restrict__ multi_count,
                      float* __restrict__ ln_mean, float* __restrict__ ln_rstd, float eps) {
    const int lane = threadIdx.x & 31;
    const int n_out = row_off[B];
    const int A2 = lay[TASU_GL_A2], A4 = lay[TASU_GL_A4], AX = lay[TASU_GL_AX], O4 = lay[TASU_GL_O4], OX = lay[TASU_GL_OX];
    const int NS = lay[TASU_GL_NS], N2 = lay[TASU_GL_N2], N4 = lay[TASU_GL_N4];
    const int fill1 = A2 - NS, fill2 = A4 - (A2 + 2 * N2), fill3 = AX - (A4 + 4 * N4);
    const int n_items = n_out + fill1 + fill2 + fill3;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const bool vec = (K % 8 == 0) && ((ldx % 8) == 0) && ((ldg % 8) == 0);
    const float mean_p = 1.f / (float)V;
    // perm entries beyond the live candidates: -1 = "zero row" for the row gather that follows the projector
    for (int64_t i = n_out + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < max_out; i += (int64_t)gridDim.x * blockDim.x)
        perm[i] = -1;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_items; r += warps) {
        if (r >= n_out) {
            // filler rows between the regions: zero rows of weight 0 (whole groups of them pool to zero rows)
            int i = r - n_out, a;
            if (i < fill1) a = NS + i;
            else if (i < fill1 + fill2) a = A2 + 2 * N2 + (i - fill1);
            else a = A4 + 4 * N4 + (i - fill1 - fill2);
            if (a < max_a) {
                zero_row_bf16(xg + (int64_t)a * ldg, K, vec, lane);
                if (lane == 0) { g_max[a] = 0.f; g_inv[a] = 0.f; }
            }
            if (i < fill1 && a < max_o && lane == 0 && ln_mean != nullptr) { ln_mean[a] = mean_p; ln_rstd[a] = 0.f; }
            continue;
        }
        int lo = 0, hi = B;                                    // utterance of packed row r
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (row_off[mid] <= r) lo = mid; else hi = mid; }
        const int b = lo, j = r - row_off[b];
        const int64_t pj = (int64_t)b * T + j;
        const int t0 = seg_start[pj], n = seg_len[pj];
        const int cls = run_class(n), s = slot[pj];
        int a0, out, rows_here, tail0 = 0;                     // first A row, pooled row, A rows of the group
        float weight;
        if (cls == 1) { const int k = base[B + b] + s; a0 = A2 + 2 * k; out = A2 + k; rows_here = 2; weight = 0.5f; }
        else if (cls == 2) { const int k = base[2 * B + b] + s; a0 = A4 + 4 * k; out = O4 + k; rows_here = 4; weight = 1.f / (float)n; }
        else { a0 = base[b] + s; out = a0; rows_here = 1; weight = 1.f; tail0 = base[3 * B + b] + xoff[pj]; }
        if (lane == 0) {
            if (r < max_out) perm[r] = out < max_proj ? out : -1;   // beyond the projector's row capacity: batch is redone
            if (cls == 3 && out < max_o && r < max_out) {                     // the work list has max_out slots
                pk_len[out] = n;
                tail_src[out] = OX + tail0;
                multi_rows[atomicAdd(multi_count, 1)] = out;                   // work list of pool_tail (any order)
            }
        }
        const int n_rows = cls == 3 ? n : rows_here;
        for (int f = 0; f < n_rows; ++f) {
            const int64_t a = (cls == 3 && f > 0) ? (int64_t)AX + tail0 + (f - 1) : (int64_t)a0 + f;
            if (a >= max_a) continue;
            __nv_bfloat16* dst = xg + a * ldg;
            if (f >= n) {                                                          // the zero row of a 3-frame run
                zero_row_bf16(dst, K, vec, lane);
                if (lane == 0) { g_max[a] = 0.f; g_inv[a] = 0.f; }
                continue;
            }
            copy_row_bf16(x + ((int64_t)b * (T + n_prefix) + n_prefix + t0 + f) * ldx, dst, K, vec, lane);
            if (lane == 0) {
                const int64_t fr = (int64_t)b * T + t0 + f;
                const float sx = row_sumexp[fr];
                g_max[a] = row_max[fr];
                g_inv[a] = weight / sx;
                if (f == 0 && rows_here == 1 && out < max_o && ln_mean != nullptr) {
                    // single-frame row: mean p = 1/V, sum p^2 = s2/s^2 (rows of long runs are overwritten by pool_tail)
                    const float q = row_sumexp2 ? row_sumexp2[fr] / (sx * sx) : 0.f;
                    float var = q / (float)V - mean_p * mean_p;
                    var = var < 0.f ? 0.f : var;
                    ln_mean[out] = mean_p;
                    ln_rstd[out] = rsqrtf(var + eps);
                }
            }
        }
    }
}

// LayerNorm statistics of the pooled rows of the G2 / G4 regions: sum p^2 from the per-(column tile, epilogue group)
// partial sums the GEMM epilogue left in q_part [n_parts][ldq], added in a FIXED order: a CTA takes 32 rows, its 8 warps
// each add every 8th partial of a row (independent, coalesced loads), the 8 sub-sums meet in shared memory.
__global__ void __launch_bounds__(256)
group_ln_finish_kernel(const float* __restrict__ q_part, int64_t ldq, int n_parts, const int32_t* __restrict__ lay, int V,
                       int64_t max_o, float* __restrict__ ln_mean, float* __restrict__ ln_rstd, float eps) {
    __shared__ float sub[8][33];
    const int A2 = lay[TASU_GL_A2], OX = lay[TASU_GL_OX];
    const int64_t n_rows = min((int64_t)(OX - A2), ldq);
    const int rx = threadIdx.x & 31, ky = threadIdx.x >> 5;
    const float mean = 1.f / (float)V;
    for (int64_t r0 = (int64_t)blockIdx.x * 32; r0 < n_rows; r0 += (int64_t)gridDim.x * 32) {
        const int64_t i = r0 + rx;
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        if (i < n_rows) {
            int k = ky;
            for (; k + 24 < n_parts; k += 32) {
                q0 += q_part[(int64_t)k * ldq + i];
                q1 += q_part[(int64_t)(k + 8) * ldq + i];
                q2 += q_part[(int64_t)(k + 16) * ldq + i];
                q3 += q_part[(int64_t)(k + 24) * ldq + i];
            }
            for (; k < n_parts; k += 8) q0 += q_part[(int64_t)k * ldq + i];
        }
        sub[ky][rx] = (q0 + q1) + (q2 + q3);
        __syncthreads();
        if (ky == 0 && i < n_rows && A2 + i < max_o) {
            float q = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) q += sub[w][rx];
            float var = q / (float)V - mean * mean;
            var = var < 0.f ? 0.f : var;
            ln_mean[A2 + i] = mean;
            ln_rstd[A2 + i] = rsqrtf(var + eps);
        }
        __syncthreads();
    }
}

}  // namespace tasu

using namespace tasu;

extern "C" int tasu_group_plan(const int32_t* seg_len, const int64_t* new_lens, int B, int T, int32_t* slot, int32_t* xoff,
                               int32_t* cnt, int32_t* base, int32_t* lay, int32_t* ticket, void* stream) {
    TASU_CHECK_ARG(B > 0 && T >= 0, "B > 0, T >= 0");
    TASU_CHECK_ARG(seg_len && new_lens && slot && xoff && cnt && base && lay && ticket, "null pointer");
    GroupScanArgs sc{ticket, cnt, base, lay};
    group_plan_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(seg_len, new_lens, T, slot, xoff, sc);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_gather_kept_rows_grouped(const void* x_bf16, int64_t ldx, int B, int T, int n_prefix, int K, int V,
                                             const int32_t* seg_start, const int32_t* seg_len, const int32_t* row_off,
                                             const int32_t* slot, const int32_t* xoff, const int32_t* base, const int32_t* lay,
                                             const float* row_max, const float* row_sumexp, const float* row_sumexp2,
                                             int64_t max_a, int64_t max_o, int64_t max_out, int64_t max_proj, void* xg_bf16,
                                             int64_t ldg,
                                             float* g_max, float* g_inv, int32_t* perm, int32_t* pk_len, int32_t* tail_src,
                                             int32_t* multi_rows, int32_t* multi_count, float* ln_mean, float* ln_rstd,
                                             float ln_eps, void* stream) {
    TASU_CHECK_ARG(B > 0 && T >= 0 && n_prefix >= 0 && K > 0 && V > 0 && ldx >= K && ldg >= K, "shape");
    TASU_CHECK_ARG(max_a > 0 && max_o > 0 && max_out > 0, "capacities");
    TASU_CHECK_ARG(x_bf16 && seg_start && seg_len && row_off && slot && xoff && base && lay && row_max && row_sumexp &&
                   xg_bf16 && g_max && g_inv && perm && pk_len && tail_src && multi_rows && multi_count, "null pointer");
    TASU_CHECK_ARG((ln_mean == nullptr) == (ln_rstd == nullptr), "ln stats come in pairs");
    TASU_CHECK_ARG(((uintptr_t)x_bf16 % 16 == 0) && ((uintptr_t)xg_bf16 % 16 == 0), "16-byte alignment");
    cudaStream_t st = (cudaStream_t)stream;
    TASU_CHECK_CUDA(cudaMemsetAsync(multi_count, 0, sizeof(int32_t), st));
    const unsigned grid = (unsigned)(tasu::sm_count() * 8);
    gather_grouped_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x_bf16, ldx, B, T, n_prefix, K, V, seg_start, seg_len, row_off,
                                                slot, xoff, base, lay, row_max, row_sumexp, row_sumexp2, max_a, max_o, max_out, max_proj,
                                                (__nv_bfloat16*)xg_bf16, ldg, g_max, g_inv, perm, pk_len, tail_src, multi_rows,
                                                multi_count, ln_mean, ln_rstd, ln_eps);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_group_ln_finish(const float* q_part, int64_t ldq, int n_parts, const int32_t* lay, int V, int64_t max_o,
                                    float* ln_mean, float* ln_rstd, float ln_eps, void* stream) {
    TASU_CHECK_ARG(ldq > 0 && n_parts > 0 && V > 0 && max_o > 0, "shape");
    TASU_CHECK_ARG(q_part && lay && ln_mean && ln_rstd, "null pointer");
    int64_t g = (ldq + 31) / 32, gmax = (int64_t)tasu::sm_count() * 4;     // 32 rows per CTA; the live row count is on the device
    if (g > gmax) g = gmax;
    group_ln_finish_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(q_part, ldq, n_parts, lay, V, max_o, ln_mean, ln_rstd, ln_eps);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
