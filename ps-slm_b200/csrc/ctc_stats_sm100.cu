// Steps 1b + 2a fused — CTC head with the softmax statistics computed in the tcgen05 GEMM epilogue (tasu_ctc_head_stats):
// logits = X·W^T + b live only in TMEM.  Replaces ctc_lo + softmax (Multitask/model/ps-slm.py:450-451, :581-582), the
// global max of :256 and the argmax of :265 without materialising the [B, T+4, 25055] tensor.  Shares the PTX wrappers,
// tile constants and tensor-map helpers of the GEMM (gemm_common.cuh).
#include "gemm_common.cuh"

namespace tasu {
namespace gemm {

// ------------------------------------------------------------------ fused CTC head + softmax statistics
// logits = X·W^T + b are produced tile by tile in TMEM and consumed in place: every epilogue thread owns
// one frame (TMEM lane) and keeps that frame's running max / sum-exp / argmax / blank logit in registers
// while its CTA sweeps a contiguous range of vocabulary tiles.  Nothing of the [frames, 25055] logits
// tensor ever reaches HBM (replaces ps-slm.py:450-451 / :581-582 + the argmax of :265).
// Work item = (m_tile, vocab split); partial statistics of the S splits are merged by ctc_stats_combine_kernel.
struct StatsParams {
    int M, N, K;              // frames (raw rows incl. prefix), vocab, encoder width
    int splits, nt_per;       // vocab splits and n-tiles per split
    int blank;
    const float* bias;
    float* part_max; float* part_sum; float* part_sum2; int32_t* part_arg; float* xb_raw;
};

constexpr int kStatsSmemBytes = kStages * kStageBytes + 2 * BN * 4 + 128;
constexpr int kStatsSmemBytesARes = 8 * kABytes + 3 * kBBytes + 2 * BN * 4 + 128;    // resident A + 3-stage B ring
constexpr int kStatsThreads = 384;                   // warps 0-3 control, warps 4-11 epilogue
constexpr int kStatsEpiThreads = 256;

// kARes (K <= 512): the 128-frame A tile stays resident in shared memory for the whole vocabulary sweep of an item
// and only the weight tiles stream through a 3-stage ring — a third less L2→SM traffic per MMA.
// Every epilogue thread fetches its bias value of the NEXT vocabulary tile while the current one is reduced (the load
// at the top of each tile is exposed L2 latency otherwise; measured −2 % on the headline batch, profiles/r02a_ab.md).
// Variants measured and removed in round 2 (profiles/r02a_ab.md): CTA pairs (cta_group::2, 256 frames per item) +10 %
// time, 16 epilogue warps on 16-column TMEM slabs +1 %.
template <bool kARes>
__global__ void __launch_bounds__(kStatsThreads, 1)
ctc_stats_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const StatsParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    constexpr int kSt = kARes ? 3 : kStages;                                   // ring stages
    constexpr int kRingStage = kARes ? kBBytes : kStageBytes;                  // bytes per ring stage
    constexpr int kRingOff = kARes ? 8 * kABytes : 0;                          // resident A: 8 k-blocks x 16 KB
    constexpr int kTileM = BM;                                                 // frames per work item
    uint8_t* ring = smem + kRingOff;
    float* s_bias = reinterpret_cast<float*>(ring + kSt * kRingStage);         // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kSt * kRingStage + 2 * BN * 4);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kSt;
    uint64_t* tmem_full = bars + 2 * kSt;
    uint64_t* tmem_empty = bars + 2 * kSt + kAccStages;
    uint64_t* a_full = bars + 2 * kSt + 2 * kAccStages;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(a_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + kTileM - 1) / kTileM, n_tiles = (p.N + BN - 1) / BN;
    const int num_items = m_tiles * p.splits;
    const int k_blocks = (p.K + BK - 1) / BK;
#define TASU_ITEM_LOOP for (int item = blockIdx.x; item < num_items; item += gridDim.x)
#define TASU_ITEM_M0 ((item / p.splits) * kTileM)

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kSt; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kStatsEpiThreads); }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_base_slot)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, a_phase = 0;
            TASU_ITEM_LOOP {
                const int m0 = TASU_ITEM_M0;
                const int nb = (item % p.splits) * p.nt_per, ne = min(nb + p.nt_per, n_tiles);
                if (kARes) {
                    mbar_wait(a_empty, a_phase ^ 1);               // MMAs of the previous item are done with A
                    mbar_expect_tx(a_full, (uint32_t)(k_blocks * kABytes));
                    for (int kb = 0; kb < k_blocks; ++kb) tma_load_2d(&tmap_a, a_full, smem + kb * kABytes, kb * BK, m0);
                    a_phase ^= 1;
                }
                for (int nt = nb; nt < ne; ++nt) {
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sr = ring + stage * kRingStage;
                        mbar_expect_tx(&full_bar[stage], (uint32_t)kRingStage);
                        if (kARes) {
                            tma_load_2d(&tmap_b, &full_bar[stage], sr, kb * BK, nt * BN);
                        } else {
                            tma_load_2d(&tmap_a, &full_bar[stage], sr, kb * BK, m0);
                            tma_load_2d(&tmap_b, &full_bar[stage], sr + kABytes, kb * BK, nt * BN);
                        }
                        if (++stage == kSt) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, a_phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            TASU_ITEM_LOOP {
                const int nb = (item % p.splits) * p.nt_per, ne = min(nb + p.nt_per, n_tiles);
                if (kARes) { mbar_wait(a_full, a_phase); tc_fence_after(); a_phase ^= 1; }
                for (int nt = nb; nt < ne; ++nt) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sr = smem_u32(ring + stage * kRingStage);
                        const uint64_t adesc = make_smem_desc(kARes ? smem_u32(smem + kb * kABytes) : sr);
                        const uint64_t bdesc = make_smem_desc(kARes ? sr : sr + kABytes);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDesc,
                                      (kb > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[stage]);
                        if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
                        if (++stage == kSt) { stage = 0; phase ^= 1; }
                    }
                    if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
                }
                if (kARes) umma_commit(a_empty);                   // every MMA of this item has retired → A may be replaced
            }
        }
    } else if (warp >= 4) {
        // 8 epilogue warps: warp pair (q, q+4) shares TMEM lane quadrant q and splits the 256 columns in halves,
        // so the per-element softmax math keeps up with the tensor pipe (K is only 512 deep)
        const int ew = (warp - 4) & 3, half = (warp - 4) >> 2;
        const int et = threadIdx.x - 128;                          // 0..255
        constexpr float kL2e = 1.4426950408889634f;
        int acc = 0; uint32_t acc_phase = 0;
        int bbuf = 0;
        // this thread's bias value (column et) of the tile after the current one, in program order
#define TASU_BIAS_OF(n0_) (((n0_) + et) < p.N ? (p.bias ? __ldg(p.bias + (n0_) + et) : 0.f) : -INFINITY)
        float pf_bias = 0.f;
        if ((int)blockIdx.x < num_items) pf_bias = TASU_BIAS_OF(((int)blockIdx.x % p.splits) * p.nt_per * BN);
        TASU_ITEM_LOOP {
            const int split = item % p.splits;
            const int row = TASU_ITEM_M0 + ew * 32 + lane;
            const int nb = split * p.nt_per, ne = min(nb + p.nt_per, n_tiles);
            float rm = -INFINITY, rs = 0.f, rs2 = 0.f, xb = 0.f;
            int best = 0x7fffffff;
            for (int nt = nb; nt < ne; ++nt) {
                const int n0 = nt * BN;
                float* sb = s_bias + bbuf * BN;
                sb[et] = pf_bias;                                  // -inf masks the columns beyond the vocabulary
                // next tile of this item, else the first tile of this CTA's next item
                const int nxt_item = item + (int)gridDim.x;
                if (nt + 1 < ne) pf_bias = TASU_BIAS_OF((nt + 1) * BN);
                else if (nxt_item < num_items) pf_bias = TASU_BIAS_OF((nxt_item % p.splits) * p.nt_per * BN);
                asm volatile("bar.sync 1, %0;" :: "n"(kStatsEpiThreads) : "memory");
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);

                auto process = [&](uint32_t (&v)[32], int sub) {
                    const int base = n0 + sub * 32;
                    if (base >= p.N) return;
                    const float4* b4 = reinterpret_cast<const float4*>(sb + sub * 32);
                    float x[32];
                    float cm = -INFINITY;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 bb = b4[q];
                        x[4 * q] = __uint_as_float(v[4 * q]) + bb.x;
                        x[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + bb.y;
                        x[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + bb.z;
                        x[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + bb.w;
                        cm = fmaxf(cm, fmaxf(fmaxf(x[4 * q], x[4 * q + 1]), fmaxf(x[4 * q + 2], x[4 * q + 3])));
                    }
                    if (cm > rm) {                                 // rare after the first slabs
                        const float f = ex2_approx((rm - cm) * kL2e);      // rm = -inf on the first slab: 2^-inf = 0
                        rs *= f;
                        rs2 *= f * f;
                        rm = cm;
                        int j0 = 31;
#pragma unroll
                        for (int j = 30; j >= 0; --j) j0 = (x[j] == cm) ? j : j0;
                        best = base + j0;                          // first index of the maximum (torch tie rule)
                    }
                    const float mb = rm * kL2e;
                    float acc_s = 0.f, acc_s2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float e = ex2_approx(fmaf(x[j], kL2e, -mb));   // one FFMA + one MUFU.EX2 per logit
                        acc_s += e;
                        acc_s2 = fmaf(e, e, acc_s2);           // Σ exp(2(x-m)): gives Σ p² = s2/s² for the LayerNorm fold
                    }
                    rs += acc_s;
                    rs2 += acc_s2;
                    if (p.blank >= base && p.blank < base + 32) {  // uniform branch: one slab of one tile
#pragma unroll
                        for (int j = 0; j < 32; ++j) xb = (base + j == p.blank) ? x[j] : xb;
                    }
                };

                uint32_t va[32], vb[32];
                const int sub0 = half * (BN / 64), sub1 = sub0 + BN / 64;      // this warp's 4 slabs
                tmem_ld32(t_row + (uint32_t)(sub0 * 32), va);
#pragma unroll 1
                for (int sub = sub0; sub < sub1; sub += 2) {
                    tmem_ld_wait(va);
                    tmem_ld32(t_row + (uint32_t)((sub + 1) * 32), vb);
                    process(va, sub);
                    tmem_ld_wait(vb);
                    if (sub + 2 < sub1) tmem_ld32(t_row + (uint32_t)((sub + 2) * 32), va);
                    else {
                        tc_fence_before();
                        mbar_arrive(&tmem_empty[acc]);
                    }
                    process(vb, sub + 1);
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
                bbuf ^= 1;
            }
            if (row < p.M) {
                const int64_t o = (int64_t)(split * 2 + half) * p.M + row;     // ascending column order of the partials
                p.part_max[o] = rm;
                p.part_sum[o] = rs;
                p.part_sum2[o] = rs2;
                p.part_arg[o] = best;
                // the blank column lives in exactly one (tile, half): that thread publishes its logit
                const int bt = p.blank / BN, bh = (p.blank % BN) / (BN / 2);
                if (bt >= nb && bt < ne && bh == half) p.xb_raw[row] = xb;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
#undef TASU_ITEM_LOOP
#undef TASU_ITEM_M0
#undef TASU_BIAS_OF
}

// merge the per-split partial statistics and drop the prefix frames: frame (b,t) ↔ raw row b*(T+P)+P+t
__global__ void __launch_bounds__(256)
ctc_stats_combine_kernel(StatsParams p, int B, int T, int n_prefix, int32_t* __restrict__ argmax,
                         float* __restrict__ x_blank, float* __restrict__ row_max, float* __restrict__ row_sumexp,
                         float* __restrict__ row_sumexp2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * T) return;
    const int b = (int)(i / T), t = (int)(i % T);
    const int64_t r = (int64_t)b * (T + n_prefix) + n_prefix + t;
    // every partial of the frame is loaded up front (independent loads: one L2 round trip instead of ~70 dependent ones)
    constexpr int kMaxParts = 16;                                          // 8 vocabulary splits x 2 column halves
    const int parts = 2 * p.splits;                                        // (split, column half), ascending columns
    float pm[kMaxParts], ps[kMaxParts], ps2[kMaxParts];
    int pa[kMaxParts];
#pragma unroll
    for (int s = 0; s < kMaxParts; ++s) {
        const bool on = s < parts;
        const int64_t o = (int64_t)s * p.M + r;
        pm[s] = on ? p.part_max[o] : -INFINITY;
        pa[s] = on ? p.part_arg[o] : 0x7fffffff;
        ps[s] = on ? p.part_sum[o] : 0.f;
        ps2[s] = on ? p.part_sum2[o] : 0.f;
    }
    float m = -INFINITY; int a = 0x7fffffff;
#pragma unroll
    for (int s = 0; s < kMaxParts; ++s)
        if (pm[s] > m) { m = pm[s]; a = pa[s]; }                           // ties keep the lower part = lower index
    float sum = 0.f, sum2 = 0.f;
#pragma unroll
    for (int s = 0; s < kMaxParts; ++s) {
        if (s < parts) {
            const float f = exp2f((pm[s] - m) * 1.4426950408889634f);
            sum += ps[s] * f;
            sum2 += ps2[s] * f * f;
        }
    }
    if (row_sumexp2) row_sumexp2[i] = sum2;
    argmax[i] = a;
    x_blank[i] = p.xb_raw[r];
    row_max[i] = m;
    row_sumexp[i] = sum;
}

}  // namespace gemm
}  // namespace tasu

using namespace tasu;
using namespace tasu::gemm;

static void pick_splits(int m_tiles, int n_tiles, int grid, int* splits, int* nt_per) {
    double best_eff = -1.0; int best_s = 1, best_per = n_tiles;
    for (int s = 1; s <= 8 && s <= n_tiles; ++s) {
        const int per = (n_tiles + s - 1) / s;
        const int eff_s = (n_tiles + per - 1) / per;          // splits actually used with this tile count
        const long items = (long)m_tiles * eff_s;
        const long waves = (items + grid - 1) / grid;
        const double eff = (double)items / (double)(waves * grid);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_s = eff_s; best_per = per; }
    }
    *splits = best_s; *nt_per = best_per;
}

extern "C" int64_t tasu_ctc_head_stats_workspace(int B, int T, int n_prefix) {
    const int64_t rows = (int64_t)B * (T + n_prefix);
    // 16 partials per frame (8 vocabulary splits x 2 column halves) x (max, sum, sum2, arg) + blank logits
    return rows * 16 * 16 + rows * 4 + 256;
}

extern "C" int tasu_ctc_head_stats(const void* x_bf16, int64_t ldx, const void* w_bf16, int64_t ldw, const float* bias,
                                   int B, int T, int n_prefix, int V, int K, int blank_id, int32_t* argmax,
                                   float* x_blank, float* row_max, float* row_sumexp, float* row_sumexp2,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    TASU_CHECK_ARG(B >= 0 && T >= 0 && n_prefix >= 0 && V > 0 && K > 0, "shape");
    TASU_CHECK_ARG(blank_id >= 0 && blank_id < V, "blank_id out of range");
    TASU_CHECK_ARG(ldx >= K && ldw >= K, "leading dimension too small");
    const int64_t rows64 = (int64_t)B * (T + n_prefix);
    TASU_CHECK_ARG(rows64 < (1LL << 31), "too many frames for one call");
    if ((int64_t)B * T == 0) return TASU_OK;
    TASU_CHECK_ARG(x_bf16 && w_bf16 && argmax && x_blank && row_max && row_sumexp && workspace, "null pointer");
    TASU_CHECK_ARG(workspace_bytes >= tasu_ctc_head_stats_workspace(B, T, n_prefix), "workspace too small");
    TASU_CHECK_ARG(((uintptr_t)x_bf16 % 16 == 0) && ((uintptr_t)w_bf16 % 16 == 0) && ((uintptr_t)workspace % 16 == 0) &&
                   (ldx * 2) % 16 == 0 && (ldw * 2) % 16 == 0, "16-byte alignment of operands");
    const int M = (int)rows64;
    CUtensorMap ma, mb;
    int rc = make_map(&ma, x_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, ldx, BM, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
    rc = make_map(&mb, w_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, V, K, ldw, BN, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
    const int m_tiles = (M + BM - 1) / BM, n_tiles = (V + BN - 1) / BN;
    int grid = sm_count();
    StatsParams p{};
    p.M = M; p.N = V; p.K = K; p.blank = blank_id; p.bias = bias;
    pick_splits(m_tiles, n_tiles, grid, &p.splits, &p.nt_per);
    if (grid > m_tiles * p.splits) grid = m_tiles * p.splits;
    const int64_t cap = 16;                                  // partial slots per frame in the workspace layout
    float* ws = reinterpret_cast<float*>(workspace);
    p.part_max = ws;
    p.part_sum = ws + cap * M;
    p.part_sum2 = ws + 2 * cap * M;
    p.part_arg = reinterpret_cast<int32_t*>(ws + 3 * cap * M);
    p.xb_raw = ws + 4 * cap * M;
    cudaStream_t st = (cudaStream_t)stream;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(ctc_stats_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatsSmemBytesARes);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(ctc_stats_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatsSmemBytes);
    });
    TASU_CHECK_CUDA(attr_err);
    if (K <= 8 * BK) ctc_stats_kernel<true><<<grid, kStatsThreads, kStatsSmemBytesARes, st>>>(ma, mb, p);
    else ctc_stats_kernel<false><<<grid, kStatsThreads, kStatsSmemBytes, st>>>(ma, mb, p);
    TASU_CHECK_LAUNCH();
    const int64_t frames = (int64_t)B * T;
    ctc_stats_combine_kernel<<<(unsigned)((frames + 255) / 256), 256, 0, st>>>(p, B, T, n_prefix, argmax, x_blank, row_max, row_sumexp, row_sumexp2);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
