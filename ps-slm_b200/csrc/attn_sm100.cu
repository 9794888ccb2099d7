// Cross-attention projector (projector.py:104-126, "cross-attention"): Z_h = softmax(Q_h K_h^T) K_h for all heads in ONE
// launch, with the probabilities never leaving the SM (tasu_attn_softmax_pv).
//
// The composed path (tasu_gemm_bf16_tn(EPI_SOFTMAX) + tasu_gemm_bf16_f32 per head) writes and re-reads a
// [rows, 151936] bf16 probability matrix per head — 2 x 3 GB of HBM traffic per head at the config-2 size.  Here a CTA
// owns a work item (128 query rows, one head) and sweeps the keys in tiles of 128:
//
//   warp 0   TMA producer: the item's Q tile once, then one table tile [128 keys x dp] per step through a ring.  Keys
//            and values are the SAME table slice, so the one shared-memory tile feeds both contractions: as the K-major
//            B operand of S = Q K^T and, through an MN-major descriptor, as the B operand of O += P K.
//   warp 1   MMA issuer (one thread): S(j+1) is issued BEFORE O += P(j) K(j), so the tensor pipe computes the next score
//            tile while the softmax warps turn the current one into probabilities.
//   warps 4-7 softmax + epilogue, one query row per thread: S from TMEM (two 128-column accumulators), p = 2^(s log2e -
//            max log2e + log2(1/sum)) with the row max / sum of a preceding statistics pass (tasu_ctc_head_stats — no
//            running maximum, no rescaling of O), bf16, written into a 128-byte-swizzled K-major shared-memory tile =
//            the A operand of the second MMA; the tile is handed over in two halves of 64 keys so that the softmax of
//            the next tile starts while the second half of the current one is still being multiplied.
//            O [128 x dp] fp32 stays in TMEM for the whole sweep and is written out once per item.
//
// Row maxima: either given (row_max / row_inv of a statistics pass: the probabilities are then exactly the bf16 values
// the composed path stores), or — row_max == NULL, the default of the host mirror — found by the kernel itself in a FIRST
// sweep over the keys that only runs S = Q K^T and a running fmax per row (no exponentials: the sweep is bound by the
// tensor pipe / the L2 feed, where the stand-alone statistics pass is bound by MUFU.EX2); the second sweep then uses
// unnormalised p = 2^((s - max) log2e), sums them per row in fp32 and the epilogue divides O by the sum.
#include "gemm_common.cuh"

namespace tasu {
namespace gemm {

constexpr int kAtKeys = 128;                       // keys per tile = N of the score MMA = K of the output MMA
constexpr int kAtBox = 128 * 64 * 2;               // one TMA box [128 rows][64 columns] bf16, 128-byte swizzle: 16 KB

template <int DP> struct AttnCfg {
    static constexpr int kbd = DP / 64;                                  // 64-column boxes per tile
    static constexpr int kTStages = DP <= 192 ? 3 : 2;                   // table-tile ring (227 KB of shared memory)
    static constexpr int kQBytes = kbd * kAtBox, kTBytes = kbd * kAtBox, kPBytes = 2 * kAtBox;
    static constexpr int kSmem = kQBytes + kTStages * kTBytes + kPBytes + 256;
    static_assert(kSmem <= 227 * 1024, "shared memory exceeds the 227 KB a CTA can opt into");
    // kind::f16 descriptors (gemm_common.cuh): S = Q K^T is 128 x 128, both operands K-major; O += P K is 128 x DP with
    // the B operand (the table tile read as [keys, d]) MN-major
    static constexpr uint32_t kIdescS = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kAtKeys >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    static constexpr uint32_t kIdescO = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(DP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

struct AttnParams {
    int N, V2, heads;
    const float* row_max;          // [heads, stat_stride]: max_k s[row, k]
    const float* row_inv;          // [heads, stat_stride]: 1 / sum_k exp(s - max)
    int64_t stat_stride;
    float* Z;                      // [N, ldz] fp32, head h in columns [h DP, h DP + DP)
    int64_t ldz;
    // key split (row_max == NULL only): an item is (128 query rows, head, split s) and sweeps the key tiles
    // [s J / splits, (s + 1) J / splits); it leaves its UNNORMALISED output (relative to its own row maxima), the maxima
    // and the sums of its probabilities in the workspace and attn_merge_kernel combines the splits.  splits = 1: Z directly.
    int splits;
    float* ws_o;                   // [items][128][DP] fp32
    float* ws_ml;                  // [items][128][2]: row maximum (raw score), sum of the unnormalised probabilities
};

// MN-major operand in [128 K-rows][64 MN] boxes: 8-row K groups 1024 B apart (SBO), the next 64 MN elements one box
// (16 KB) further (LBO)
__device__ __forceinline__ uint64_t make_smem_desc_mn_box128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((uint32_t)kAtBox >> 4) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int DP>
__global__ void __launch_bounds__(256, 1)
attn_softmax_pv_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_t, const AttnParams p) {
    using C = AttnCfg<DP>;
    constexpr int kbd = C::kbd, kTS = C::kTStages;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* sQ = smem;
    uint8_t* sT = sQ + C::kQBytes;
    uint8_t* sP = sT + kTS * C::kTBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::kPBytes);
    uint64_t* q_full = bars;            // TMA -> MMA: the item's Q tile
    uint64_t* q_empty = bars + 1;       // MMA -> TMA: every MMA of the item has retired
    uint64_t* t_full = bars + 2;        // [kTS] TMA -> MMA
    uint64_t* t_empty = t_full + kTS;   // [kTS] MMA -> TMA (after the tile's second contraction)
    uint64_t* s_full = t_empty + kTS;   // [2] MMA -> softmax: score accumulator ready
    uint64_t* s_empty = s_full + 2;     // [2] softmax -> MMA: accumulator read out (128 arrivals)
    uint64_t* p_full = s_empty + 2;     // [2] softmax -> MMA: half h (64 keys) of the probability tile written (128 arrivals)
    uint64_t* p_empty = p_full + 2;     // [2] MMA -> softmax: half h consumed
    uint64_t* o_full = p_empty + 2;     // MMA -> epilogue: O complete
    uint64_t* o_empty = o_full + 1;     // epilogue -> MMA: O read out (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.N + 127) / 128, n_items = m_tiles * p.heads * p.splits, J = (p.V2 + kAtKeys - 1) / kAtKeys;
    // item -> (head, first query row, key-tile range): every role derives the same ranges from the same expressions
#define TASU_AT_ITEM(it) \
    const int sp_ = (it) % p.splits, bi_ = (it) / p.splits; \
    const int head = bi_ / m_tiles, m0 = (bi_ % m_tiles) * 128; \
    const int j0 = (int)((int64_t)sp_ * J / p.splits), j1 = (int)((int64_t)(sp_ + 1) * J / p.splits); \
    (void)head; (void)m0
    const bool find_max = p.row_max == nullptr;                  // first sweep: row maxima only
    const int n_sweeps = find_max ? 2 : 1;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_t); }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (int s = 0; s < kTS; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 128);
            mbar_init(&p_full[s], 128); mbar_init(&p_empty[s], 1);
        }
        mbar_init(o_full, 1); mbar_init(o_empty, 128);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;      // columns [0,128) S0, [128,256) S1, [256, 256+DP) O

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t g = 0, n = 0;                                        // running tile / item counters of this CTA
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
                TASU_AT_ITEM(it);
                const int c0 = head * DP;
                mbar_wait(q_empty, (n & 1) ^ 1);
                mbar_expect_tx(q_full, C::kQBytes);
#pragma unroll
                for (int b = 0; b < kbd; ++b) tma_load_2d(&tmap_q, q_full, sQ + b * kAtBox, c0 + 64 * b, m0);
                for (int sweep = 0; sweep < n_sweeps; ++sweep)
                    for (int j = j0; j < j1; ++j, ++g) {
                        const uint32_t st = g % kTS;
                        mbar_wait(&t_empty[st], ((g / kTS) & 1) ^ 1);
                        mbar_expect_tx(&t_full[st], C::kTBytes);
#pragma unroll
                        for (int b = 0; b < kbd; ++b)
                            tma_load_2d(&tmap_t, &t_full[st], sT + st * C::kTBytes + b * kAtBox, c0 + 64 * b, j * kAtKeys);
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t q_addr = smem_u32(sQ), t_addr = smem_u32(sT), p_addr = smem_u32(sP);
            const uint32_t d_o = tmem_base + 256u;
            // S(g) = Q K(g)^T into score accumulator g % 2
            auto issue_s = [&](uint32_t g) {
                const uint32_t st = g % kTS, sb = g & 1;
                mbar_wait(&t_full[st], (g / kTS) & 1);
                mbar_wait(&s_empty[sb], ((g >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_s = tmem_base + sb * (uint32_t)kAtKeys;
#pragma unroll
                for (int b = 0; b < kbd; ++b) {
                    const uint64_t adesc = make_smem_desc(q_addr + b * kAtBox);
                    const uint64_t bdesc = make_smem_desc(t_addr + st * C::kTBytes + b * kAtBox);
#pragma unroll
                    for (int k = 0; k < 64 / UMMA_K; ++k)
                        umma_bf16(d_s, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), C::kIdescS, (b > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&s_full[sb]);
            };
            uint32_t g = 0, gp = 0, n = 0;                               // tiles issued, tiles of second sweeps, items
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
                TASU_AT_ITEM(it);
                mbar_wait(q_full, n & 1);
                tc_fence_after();
                if (find_max)
                    for (int j = j0; j < j1; ++j, ++g) {                   // first sweep: scores only
                        issue_s(g);
                        umma_commit(&t_empty[g % kTS]);
                    }
                issue_s(g);
                for (int j = j0; j < j1; ++j, ++g, ++gp) {
                    if (j + 1 < j1) issue_s(g + 1);
                    const uint32_t st = g % kTS;
                    const uint64_t bdesc = make_smem_desc_mn_box128(t_addr + st * C::kTBytes);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {                          // the two 64-key halves of the probability tile
                        mbar_wait(&p_full[h], gp & 1);
                        tc_fence_after();
                        if (j == j0 && h == 0) { mbar_wait(o_empty, (n & 1) ^ 1); tc_fence_after(); }
                        const uint64_t adesc = make_smem_desc(p_addr + h * kAtBox);
#pragma unroll
                        for (int k = 0; k < 64 / UMMA_K; ++k)
                            umma_bf16(d_o, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(128 * (4 * h + k)), C::kIdescO,
                                      (j > j0 || h > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&p_empty[h]);
                    }
                    umma_commit(&t_empty[st]);
                }
                umma_commit(o_full);
                umma_commit(q_empty);
            }
        }
    } else if (warp >= 4) {
        // ===================== softmax + epilogue =====================
        const int ew = warp - 4, r = ew * 32 + lane;                     // TMEM lanes [32 ew, 32 ew + 32): one row per thread
        const uint32_t lane_bits = (uint32_t)(ew * 32) << 16;
        const uint32_t p_row = smem_u32(sP) + (uint32_t)(r * 128);
        const int sw = r & 7;
        uint32_t g = 0, gp = 0, n = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
            TASU_AT_ITEM(it);
            const int grow = m0 + r;
            float rowc = -INFINITY;                                      // given statistics, rows beyond N: p = 0
            float lsum = 0.f;                                            // find_max: sum of the unnormalised probabilities
            float rmax = -INFINITY;
            if (find_max) {
                for (int j = j0; j < j1; ++j, ++g) {
                    const uint32_t sb = g & 1;
                    const int n_valid = min(kAtKeys, p.V2 - j * kAtKeys);
                    mbar_wait(&s_full[sb], (g >> 1) & 1);
                    tc_fence_after();
                    const uint32_t t_row = tmem_base + lane_bits + sb * (uint32_t)kAtKeys;
                    uint32_t va[32], vb[32];
                    tmem_ld32(t_row, va);
#pragma unroll
                    for (int c = 0; c < 4; c += 2) {
                        tmem_ld_wait(va);
                        tmem_ld32(t_row + (uint32_t)((c + 1) * 32), vb);
#pragma unroll
                        for (int e = 0; e < 32; ++e) rmax = fmaxf(rmax, 32 * c + e < n_valid ? __uint_as_float(va[e]) : -INFINITY);
                        tmem_ld_wait(vb);
                        if (c == 0) tmem_ld32(t_row + 64u, va);
                        else { tc_fence_before(); mbar_arrive(&s_empty[sb]); }
#pragma unroll
                        for (int e = 0; e < 32; ++e) rmax = fmaxf(rmax, 32 * (c + 1) + e < n_valid ? __uint_as_float(vb[e]) : -INFINITY);
                    }
                }
                rowc = -rmax * kLog2e;
            } else if (grow < p.N) {
                const float mx = __ldg(p.row_max + (int64_t)head * p.stat_stride + grow);
                const float iv = __ldg(p.row_inv + (int64_t)head * p.stat_stride + grow);
                rowc = fmaf(-mx, kLog2e, __log2f(fmaxf(iv, 1e-37f)));
            }
            for (int j = j0; j < j1; ++j, ++g, ++gp) {
                const uint32_t sb = g & 1;
                const int n_valid = min(kAtKeys, p.V2 - j * kAtKeys);     // keys beyond V2 (zero rows of the tile): p = 0
                mbar_wait(&s_full[sb], (g >> 1) & 1);
                tc_fence_after();
                const uint32_t t_row = tmem_base + lane_bits + sb * (uint32_t)kAtKeys;
                uint32_t va[32], vb[32];
                tmem_ld32(t_row, va);
                auto store_chunk = [&](uint32_t (&v)[32], int c) {       // columns [32 c, 32 c + 32) of the tile
                    const uint32_t base = p_row + (uint32_t)((c >> 1) * kAtBox);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int col = 32 * c + 8 * q + e;
                            const float pv = ex2_approx(fmaf(__uint_as_float(v[8 * q + e]), kLog2e, rowc));
                            f[e] = col < n_valid ? pv : 0.f;
                            lsum += f[e];
                        }
                        st_shared_u4(base + (uint32_t)(((((c & 1) * 4 + q)) ^ sw) * 16),
                                     pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                    }
                };
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // chunk 2h in va, chunk 2h+1 into vb
                    tmem_ld_wait(va);
                    tmem_ld32(t_row + (uint32_t)((2 * h + 1) * 32), vb);
                    mbar_wait(&p_empty[h], (gp & 1) ^ 1);                 // the previous tile's half h has been multiplied
                    store_chunk(va, 2 * h);
                    tmem_ld_wait(vb);
                    if (h == 0) tmem_ld32(t_row + 64u, va);
                    else { tc_fence_before(); mbar_arrive(&s_empty[sb]); }   // the score accumulator is free again
                    store_chunk(vb, 2 * h + 1);
                    fence_proxy_async_smem();                            // generic-proxy stores -> the MMA's async-proxy reads
                    mbar_arrive(&p_full[h]);
                }
            }
            // ---- the item's output: O [128 x DP] fp32 from TMEM to Z
            mbar_wait(o_full, n & 1);
            tc_fence_after();
            const bool split = p.splits > 1;
            // key split: the unnormalised output of this key range (relative to its own row maxima) goes to the workspace
            float* zrow = split ? p.ws_o + ((int64_t)it * 128 + r) * DP : p.Z + (int64_t)grow * p.ldz + (int64_t)head * DP;
            const float zs = (find_max && !split) ? 1.f / lsum : 1.f;    // lsum >= 1: the row maximum contributes 2^0
            if (split) *reinterpret_cast<float2*>(p.ws_ml + ((int64_t)it * 128 + r) * 2) = make_float2(rmax, lsum);
#pragma unroll 1
            for (int c = 0; c < DP / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_bits + 256u + (uint32_t)(32 * c), v);
                tmem_ld_wait(v);
                if (split || grow < p.N) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(zrow + 32 * c + 4 * q) =
                            make_float4(__uint_as_float(v[4 * q]) * zs, __uint_as_float(v[4 * q + 1]) * zs,
                                        __uint_as_float(v[4 * q + 2]) * zs, __uint_as_float(v[4 * q + 3]) * zs);
                }
            }
            tc_fence_before();
            mbar_arrive(o_empty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

#undef TASU_AT_ITEM

// Combines the key splits of tasu_attn_softmax_pv: per (128-row tile, head) and row, m = max_s m_s,
// Z = sum_s 2^((m_s - m) log2e) O_s / sum_s 2^((m_s - m) log2e) l_s — in split order, deterministic.  One CTA per tile.
constexpr int kAtMaxSplits = 8;
__global__ void __launch_bounds__(256)
attn_merge_kernel(const float* __restrict__ ws_o, const float* __restrict__ ws_ml, int N, int heads, int dp, int splits,
                  float* __restrict__ Z, int64_t ldz) {
    __shared__ float w[kAtMaxSplits][128];
    const int m_tiles = (N + 127) / 128;
    const int bi = blockIdx.x, head = bi / m_tiles, m0 = (bi % m_tiles) * 128;
    const int64_t item0 = (int64_t)bi * splits;
    if (threadIdx.x < 128) {
        const int r = threadIdx.x;
        float m = -INFINITY;
        for (int s = 0; s < splits; ++s) m = fmaxf(m, ws_ml[((item0 + s) * 128 + r) * 2]);
        float l = 0.f, e[kAtMaxSplits];
        for (int s = 0; s < splits; ++s) {
            e[s] = exp2f((ws_ml[((item0 + s) * 128 + r) * 2] - m) * kLog2e);
            l = fmaf(e[s], ws_ml[((item0 + s) * 128 + r) * 2 + 1], l);
        }
        const float inv = 1.f / l;
        for (int s = 0; s < splits; ++s) w[s][r] = e[s] * inv;
    }
    __syncthreads();
    const int q4 = dp / 4;
    for (int idx = threadIdx.x; idx < 128 * q4; idx += blockDim.x) {
        const int r = idx / q4, c = (idx % q4) * 4;
        if (m0 + r >= N) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < splits; ++s) {
            const float4 o = *reinterpret_cast<const float4*>(ws_o + ((item0 + s) * 128 + r) * dp + c);
            const float ws = w[s][r];
            acc.x = fmaf(ws, o.x, acc.x); acc.y = fmaf(ws, o.y, acc.y); acc.z = fmaf(ws, o.z, acc.z); acc.w = fmaf(ws, o.w, acc.w);
        }
        *reinterpret_cast<float4*>(Z + (int64_t)(m0 + r) * ldz + (int64_t)head * dp + c) = acc;
    }
}

// HOST: the number of key splits that fills the last wave best: time ~ ceil(items S / SMs) / S item lengths, with a small
// charge per split for the workspace round trip (ties go to fewer splits); never more splits than key tiles
static int choose_attn_splits(int items, int sms, int J) {
    int best = 1;
    double best_cost = 1e30;
    for (int S = 1; S <= kAtMaxSplits && S <= J; ++S) {
        const int64_t waves = ((int64_t)items * S + sms - 1) / sms;
        const double cost = (double)waves / S * (1.0 + 0.004 * (S - 1));
        if (cost < best_cost - 1e-12) { best_cost = cost; best = S; }
    }
    return best;
}

template <int DP>
static int launch_attn(int grid, cudaStream_t st, const CUtensorMap& mq, const CUtensorMap& mt, const AttnParams& p) {
    auto kern = attn_softmax_pv_kernel<DP>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<DP>::kSmem); });
    TASU_CHECK_CUDA(attr_err);
    kern<<<grid, 256, AttnCfg<DP>::kSmem, st>>>(mq, mt, p);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

}  // namespace gemm
}  // namespace tasu

using namespace tasu;
using namespace tasu::gemm;

extern "C" int64_t tasu_attn_split_plan(int N, int V2, int heads, int dp, int* splits_host) {
    const int items = ((N + 127) / 128) * heads, J = (V2 + kAtKeys - 1) / kAtKeys;
    const int S = (N > 0 && V2 > 0 && heads > 0) ? choose_attn_splits(items, sm_count(), J) : 1;
    if (splits_host) *splits_host = S;
    return S > 1 ? (int64_t)items * S * 128 * (dp + 2) * 4 : 0;
}

extern "C" int tasu_attn_softmax_pv_ws(const void* Q_bf16, int64_t ldq, const void* table_bf16, int64_t ldt, int N, int V2,
                                       int heads, int dp, const float* row_max, const float* row_inv, int64_t stat_stride,
                                       float* Z, int64_t ldz, void* workspace, int64_t workspace_bytes, void* stream) {
    TASU_CHECK_ARG(N >= 0 && V2 > 0 && heads > 0, "shape");
    TASU_CHECK_ARG(dp == 64 || dp == 128 || dp == 192 || dp == 256, "head width must be 64, 128, 192 or 256 (see tasu_attn_softmax_pv_supported)");
    TASU_CHECK_ARG(ldq >= (int64_t)heads * dp && ldt >= (int64_t)heads * dp && ldz >= (int64_t)heads * dp &&
                   (row_max == nullptr || stat_stride >= N), "leading dimension too small");
    if (N == 0) return TASU_OK;
    TASU_CHECK_ARG(Q_bf16 && table_bf16 && Z, "null pointer");
    TASU_CHECK_ARG((row_max == nullptr) == (row_inv == nullptr), "row_max / row_inv come in pairs (both NULL: the kernel finds the maxima)");
    TASU_CHECK_ARG(((uintptr_t)Q_bf16 % 16 == 0) && ((uintptr_t)table_bf16 % 16 == 0) && ((uintptr_t)Z % 16 == 0),
                   "base pointers must be 16-byte aligned");
    TASU_CHECK_ARG((ldq * 2) % 16 == 0 && (ldt * 2) % 16 == 0 && (ldz * 4) % 16 == 0, "row pitches must be multiples of 16 bytes");
    CUtensorMap mq, mt;
    int rc = make_map(&mq, Q_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, (int64_t)heads * dp, ldq, 128, 64, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
    rc = make_map(&mt, table_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, V2, (int64_t)heads * dp, ldt, 128, 64, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
    // key split (self-contained mode with a workspace): fills the last wave of (row tile, head) items
    int splits = 1;
    if (row_max == nullptr && workspace != nullptr) {
        int want = 1;
        const int64_t need = tasu_attn_split_plan(N, V2, heads, dp, &want);
        TASU_CHECK_ARG(workspace_bytes >= need, "workspace smaller than tasu_attn_split_plan bytes");
        TASU_CHECK_ARG((uintptr_t)workspace % 16 == 0, "workspace must be 16-byte aligned");
        splits = want;
    }
    const int items = ((N + 127) / 128) * heads * splits, sms = sm_count();
    AttnParams p{N, V2, heads, row_max, row_inv, stat_stride, Z, ldz, splits, nullptr, nullptr};
    if (splits > 1) {
        p.ws_o = (float*)workspace;
        p.ws_ml = p.ws_o + (int64_t)items * 128 * dp;
    }
    const int grid = items < sms ? items : sms;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dp) {
        case 64: rc = launch_attn<64>(grid, st, mq, mt, p); break;
        case 128: rc = launch_attn<128>(grid, st, mq, mt, p); break;
        case 192: rc = launch_attn<192>(grid, st, mq, mt, p); break;
        default: rc = launch_attn<256>(grid, st, mq, mt, p); break;
    }
    if (rc != TASU_OK || splits == 1) return rc;
    attn_merge_kernel<<<((N + 127) / 128) * heads, 256, 0, st>>>(p.ws_o, p.ws_ml, N, heads, dp, splits, Z, ldz);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_attn_softmax_pv(const void* Q_bf16, int64_t ldq, const void* table_bf16, int64_t ldt, int N, int V2,
                                    int heads, int dp, const float* row_max, const float* row_inv, int64_t stat_stride,
                                    float* Z, int64_t ldz, void* stream) {
    return tasu_attn_softmax_pv_ws(Q_bf16, ldq, table_bf16, ldt, N, V2, heads, dp, row_max, row_inv, stat_stride, Z, ldz,
                                   nullptr, 0, stream);
}
