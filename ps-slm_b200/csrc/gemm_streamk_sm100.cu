// The deep-K tcgen05 GEMM of gemm_sm100.cu with a stream-K tail (tasu_gemm_bf16_tn_streamk), for one CTA per tile and —
// kPair, the default for M > 128 — for CTA pairs (tcgen05.mma.cta_group::2, one 256x256 tile per cluster of two CTAs:
// every "CTA" of the description below is then a cluster, and each of its two CTAs keeps, contributes and finishes its
// own 128 accumulator rows through its own slot and flag).
//
// Why: the projector GEMM-1 (projector.py:141, M = compressed rows of the batch, N = 2048, K = 25055) has
// ceil(M/128) x 8 output tiles — 528 for the 8341 rows of the headline batch = 3.57 waves of 148 CTAs.  With one CTA
// per tile the last wave runs at 57 % occupancy (89 % wave efficiency, profiles/r01h_ncu_full.md) and the row count is
// data dependent, so no fixed tile shape avoids it.
//
// How: the floor(tiles/grid) full waves run exactly as before (one tile per CTA, round-robin).  The `rem` tiles of the
// ragged last wave are cut along K: their rem * k_blocks K-blocks are dealt out evenly, in order, to
// min(grid, 4 * rem) CTAs, so every CTA gets a contiguous range of at most k_blocks K-blocks that touches at most two
// tiles.  A piece that ends at its tile's last K-block FINISHES the tile: it adds the partial accumulators of the
// pieces that cover the start of the tile (always CTAs with a LOWER index, written to a per-CTA fp32 slot in global
// memory and published with a release flag) in a fixed order and runs the normal epilogue.  A CTA computes the piece
// it contributes BEFORE the piece it finishes, so no CTA ever waits for a CTA that waits (no chains), and the kernel
// is launched cooperatively (all CTAs co-resident).  Sums are formed in a fixed order: results are deterministic, but
// the split tiles differ from the unsplit kernel in the last fp32 bits (different association of the K sum).
#include "gemm_common.cuh"

namespace tasu {
namespace gemm {

// ------------------------------------------------------------------ schedule (host + device; tested on the CPU)
constexpr int kSkMaxSplit = 4;                 // a tile of the tail is cut into at most this many pieces (+1 by rounding)
enum { SK_FULL = 0, SK_CONTRIB = 1, SK_FINISH = 2 };

struct SkPiece {
    int tile, kb0, kb1, kind;
    int n_contrib;                             // SK_FINISH: pieces to add, owned by CTAs cta-1, cta-2, ... cta-n_contrib
};

__host__ __device__ inline void sk_split(int num_tiles, int k_blocks, int grid, int* dp_tiles, int* rem, int* sk_ctas) {
    *dp_tiles = (num_tiles / grid) * grid;
    *rem = num_tiles - *dp_tiles;
    // a last wave that is at least 3/4 full runs as it is: cutting it costs more (partial tiles through global memory)
    // than its idle quarter (measured: 16384 x 2048 x 25055, last wave 92 % full, +1 % with the cut; profiles/r02n_gemm_ab.md)
    if (4 * *rem >= 3 * grid) { *dp_tiles = num_tiles; *rem = 0; }
    // at least one K-block per participating CTA: no CTA of [0, sk_ctas) has an empty range
    const int want = *rem * (k_blocks < kSkMaxSplit ? k_blocks : kSkMaxSplit);
    *sk_ctas = want < grid ? want : grid;
}
__host__ __device__ inline void sk_range(int64_t units, int sk_ctas, int cta, int64_t* b, int64_t* e) {
    *b = units * cta / sk_ctas;
    *e = units * (cta + 1) / sk_ctas;
}
__host__ __device__ inline int sk_kind(int kb0, int kb1, int k_blocks) {
    return kb1 == k_blocks ? (kb0 == 0 ? SK_FULL : SK_FINISH) : SK_CONTRIB;
}
// stream-K pieces of CTA `cta` in processing order (the contributed piece first); returns their number (0..2)
__host__ __device__ inline int sk_pieces(int num_tiles, int k_blocks, int grid, int cta, SkPiece out[2]) {
    int dp_tiles, rem, sk_ctas;
    sk_split(num_tiles, k_blocks, grid, &dp_tiles, &rem, &sk_ctas);
    out[0] = SkPiece{0, 0, 0, SK_FULL, 0};
    out[1] = out[0];
    if (rem == 0 || cta >= sk_ctas) return 0;
    const int64_t units = (int64_t)rem * k_blocks;
    int64_t b, e;
    sk_range(units, sk_ctas, cta, &b, &e);
    if (e <= b) return 0;
    const int t_a = (int)(b / k_blocks);
    const int64_t end_a = (int64_t)(t_a + 1) * k_blocks;
    SkPiece a, c;
    a.tile = dp_tiles + t_a;
    a.kb0 = (int)(b - (int64_t)t_a * k_blocks);
    a.kb1 = (int)((e < end_a ? e : end_a) - (int64_t)t_a * k_blocks);
    a.kind = sk_kind(a.kb0, a.kb1, k_blocks);
    a.n_contrib = 0;
    if (a.kind == SK_FINISH) {                 // CTAs below this one whose range reaches into the tile
        const int64_t t_begin = (int64_t)t_a * k_blocks;
        for (int o = cta - 1; o >= 0; --o) {
            int64_t ob, oe;
            sk_range(units, sk_ctas, o, &ob, &oe);
            if (oe <= t_begin) break;
            ++a.n_contrib;                     // ranges are never empty (sk_split), so CTA o owns a piece of this tile
            if (ob <= t_begin) break;
        }
    }
    if (e <= end_a) { out[0] = a; return 1; }
    c.tile = dp_tiles + t_a + 1;
    c.kb0 = 0;
    c.kb1 = (int)(e - end_a);
    c.kind = sk_kind(c.kb0, c.kb1, k_blocks);   // SK_CONTRIB (SK_FULL only if piece a were empty, which it is not)
    c.n_contrib = 0;
    out[0] = c; out[1] = a;                     // contribute first, finish second
    return 2;
}

// ------------------------------------------------------------------ kernel
constexpr int kSkSlotFloats = BM * BN;          // one partial accumulator tile: 128 KB of fp32 per CTA

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Same pipeline as gemm_bf16_tn_kernel<kOutBf16, kEpi, 4, 1> (TMA producer warp, single-thread MMA issuer, two TMEM
// accumulators, 4 epilogue warps with swizzled staging + TMA stores); work items are (tile, K-block range) pieces.
template <bool kOutBf16, int kEpi, bool kPair>
__global__ void __launch_bounds__(256, 1)
gemm_streamk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const Params p, float* __restrict__ ws_slots,
                    uint32_t* __restrict__ ws_flags) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    constexpr int kStages = kPair ? kPairStages : tasu::gemm::kStages;             // shadows the one-CTA constant
    constexpr int kStageBytes = kPair ? kPairStageBytes : tasu::gemm::kStageBytes;
    constexpr int kTileM = kPair ? 2 * BM : BM;
    uint8_t* staging = smem + kStages * kStageBytes;
    float* s_aux = reinterpret_cast<float*>(staging + 2 * kStagingBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes + kAuxBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tmem_full = bars + 2 * kStages;
    uint64_t* tmem_empty = bars + 2 * kStages + kAccStages;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M_live = p.m_dev != nullptr ? min(max(__ldg(p.m_dev), 0), p.M) : p.M;
    const int m_tiles = (M_live + kTileM - 1) / kTileM, n_tiles = (p.N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int k_blocks = (p.K + BK - 1) / BK;
    // scheduling unit: a CTA, or (pair mode) a cluster of two CTAs that share every piece
    const int grid = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x, cta = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const int my_slot = (int)blockIdx.x;                                   // partial-accumulator slot / flag of THIS CTA
    auto slot_of = [&](int unit) { return kPair ? 2 * unit + (int)rank : unit; };   // same-rank CTA of another unit

    // work items of this CTA: its tiles of the full waves, then its stream-K pieces (every role walks the same list)
    int dp_tiles, rem, sk_ctas;
    sk_split(num_tiles, k_blocks, grid, &dp_tiles, &rem, &sk_ctas);
    const int n_dp = cta < dp_tiles ? (dp_tiles - cta + grid - 1) / grid : 0;
    SkPiece pieces[2];
    const int n_sk = sk_pieces(num_tiles, k_blocks, grid, cta, pieces);
    const SkPiece piece0 = pieces[0], piece1 = pieces[1];          // constant indices: the array stays in registers
    const int n_items = n_dp + n_sk;
    auto item = [&](int i) -> SkPiece {
        if (i < n_dp) { SkPiece f; f.tile = cta + i * grid; f.kb0 = 0; f.kb1 = k_blocks; f.kind = SK_FULL; f.n_contrib = 0; return f; }
        return i == n_dp ? piece0 : piece1;
    };

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); prefetch_tmap(&tmap_c); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpiThreads * (kPair ? 2 : 1)); }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_u32(tmem_base_slot)), "r"((uint32_t)kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_u32(tmem_base_slot)), "r"((uint32_t)kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (kPair) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int i = 0; i < n_items; ++i) {
                const SkPiece it = item(i);
                const int m0 = (it.tile / n_tiles) * kTileM + (kPair ? (int)rank * BM : 0), n0 = (it.tile % n_tiles) * BN;
                for (int kb = it.kb0; kb < it.kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kStageBytes;
                    if (kPair) {
                        // both CTAs' halves of the stage complete on the rank-0 barrier the MMA thread waits on
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
                        const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        tma_load_2d_pair(&tmap_a, fb, sa, kb * BK, m0);
                        tma_load_2d_pair(&tmap_b, fb, sa + kABytes, kb * BK, n0 + (int)rank * (BN / 2));
                    } else {
                        mbar_expect_tx(&full_bar[stage], kStageBytes);
                        tma_load_2d(&tmap_a, &full_bar[stage], sa, kb * BK, m0);
                        tma_load_2d(&tmap_b, &full_bar[stage], sa + kABytes, kb * BK, n0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && (!kPair || rank == 0)) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int i = 0; i < n_items; ++i) {
                const SkPiece it = item(i);
                if (kPair) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
                else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = it.kb0; kb < it.kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (kPair) umma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDescPair,
                                                  (kb > it.kb0 || k > 0) ? 1u : 0u);
                        else umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDesc,
                                       (kb > it.kb0 || k > 0) ? 1u : 0u);
                    }
                    if (kPair) {
                        umma_commit_pair(&empty_bar[stage]);
                        if (kb == it.kb1 - 1) umma_commit_pair(&tmem_full[acc]);
                    } else {
                        umma_commit(&empty_bar[stage]);
                        if (kb == it.kb1 - 1) umma_commit(&tmem_full[acc]);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;                                   // TMEM lanes [32*ew, 32*ew+32)
        const int et = ew * 32 + lane;                             // 0..127 = row of the tile
        float* s_bias = s_aux;
        float* s_colsum = s_bias + BN;
        constexpr int kSubs = BN / 32;
        constexpr int kSubsPerChunk = kOutBf16 ? 2 : 1;
        constexpr int kColsPerChunk = 32 * kSubsPerChunk;
        const uint32_t stg_u32 = smem_u32(staging);
        const int sw = et & 7;
        int acc = 0; uint32_t acc_phase = 0;
        int sbuf = 0;
        for (int i = 0; i < n_items; ++i) {
            const SkPiece it = item(i);
            const int m0 = (it.tile / n_tiles) * kTileM + (kPair ? (int)rank * BM : 0), n0 = (it.tile % n_tiles) * BN;
            const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
            auto release_acc = [&]() {
                tc_fence_before();
                if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
                else mbar_arrive(&tmem_empty[acc]);
            };

            if (it.kind == SK_CONTRIB) {
                // partial accumulators → this CTA's slot, [column][row] so that a warp writes 128 contiguous bytes
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                float* slot = ws_slots + (size_t)my_slot * kSkSlotFloats + et;
                uint32_t va[32], vb[32];
                tmem_ld32(t_row, va);
#pragma unroll 1
                for (int sub = 0; sub < kSubs; sub += 2) {
                    tmem_ld_wait(va);
                    tmem_ld32(t_row + (uint32_t)((sub + 1) * 32), vb);
#pragma unroll
                    for (int j = 0; j < 32; ++j) __stcg(slot + (size_t)(sub * 32 + j) * BM, __uint_as_float(va[j]));
                    tmem_ld_wait(vb);
                    if (sub + 2 < kSubs) tmem_ld32(t_row + (uint32_t)((sub + 2) * 32), va);
                    else release_acc();
#pragma unroll
                    for (int j = 0; j < 32; ++j) __stcg(slot + (size_t)((sub + 1) * 32 + j) * BM, __uint_as_float(vb[j]));
                }
                __threadfence();                                   // this thread's slot writes before the flag
                asm volatile("bar.sync 1, %0;" :: "n"(kEpiThreads) : "memory");
                if (et == 0) st_release_gpu_u32(ws_flags + my_slot, 1u);
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
                continue;
            }

            const int grow = m0 + et;
            if (kEpi != TASU_EPI_NONE) {
                for (int c = et; c < BN; c += kEpiThreads) {
                    const int col = n0 + c;
                    s_bias[c] = col < p.N ? __ldg(p.bias + col) * (kEpi == TASU_EPI_SOFTMAX ? kLog2e : 1.f) : 0.f;
                    if (kEpi == TASU_EPI_LNFOLD_SILU || kEpi == TASU_EPI_LNFOLD) s_colsum[c] = col < p.N ? __ldg(p.colsum + col) : 0.f;
                }
            }
            float rstd = 1.f, nmean = 0.f;
            if ((kEpi == TASU_EPI_LNFOLD_SILU || kEpi == TASU_EPI_LNFOLD || kEpi == TASU_EPI_SOFTMAX) && grow < M_live) {
                rstd = __ldg(p.row_rstd + grow); nmean = -__ldg(p.row_mean + grow);
            }
            const float rowc = kEpi == TASU_EPI_SOFTMAX ? fmaf(nmean, kLog2e, __log2f(fmaxf(rstd, 1e-37f))) : 0.f;
            const int n_chunks = min(BN / kColsPerChunk, (p.N - n0 + kColsPerChunk - 1) / kColsPerChunk);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            // SK_FINISH: the pieces that cover the start of this tile were computed first by CTAs cta-1 .. cta-n_contrib
            const int n_contrib = it.kind == SK_FINISH ? it.n_contrib : 0;
            for (int c = 1; c <= n_contrib; ++c)
                while (ld_acquire_gpu_u32(ws_flags + slot_of(cta - c)) == 0u) { }

            auto process = [&](uint32_t (&v)[32], int sub) {
                const int ch = sub / kSubsPerChunk, h = sub % kSubsPerChunk;
                if (ch >= n_chunks) return;
                const uint32_t srow = stg_u32 + (uint32_t)(sbuf * kStagingBytes + et * 128);
                if (h == 0) {
                    if (et == 0) tma_store_wait_read<1>();
                    asm volatile("bar.sync 1, %0;" :: "n"(kEpiThreads) : "memory");
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                for (int c = 1; c <= n_contrib; ++c) {             // fixed order: nearest CTA first
                    const float* slot = ws_slots + (size_t)slot_of(cta - c) * kSkSlotFloats + (size_t)(sub * 32) * BM + et;
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __ldcg(slot + (size_t)j * BM);
                }
                const float4* b4 = reinterpret_cast<const float4*>(s_bias + sub * 32);
                const float4* c4 = reinterpret_cast<const float4*>(s_colsum + sub * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float x[4] = {f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]};
                    if (kEpi != TASU_EPI_NONE) {
                        const float4 bb = b4[q];
                        const float b[4] = {bb.x, bb.y, bb.z, bb.w};
                        if (kEpi == TASU_EPI_LNFOLD_SILU || kEpi == TASU_EPI_LNFOLD) {
                            const float4 cc = c4[q];
                            const float c[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                x[e] = fmaf(rstd, fmaf(nmean, c[e], x[e]), b[e]);
                                if (kEpi == TASU_EPI_LNFOLD_SILU) x[e] = silu_f(x[e]);
                            }
                        } else if (kEpi == TASU_EPI_SOFTMAX) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) x[e] = ex2_approx(fmaf(x[e], kLog2e, b[e] + rowc));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                x[e] += b[e];
                                if (kEpi == TASU_EPI_BIAS_SILU) x[e] = silu_f(x[e]);
                                if (kEpi == TASU_EPI_BIAS_RELU) x[e] = fmaxf(x[e], 0.f);
                            }
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) f[4 * q + e] = x[e];
                }
                if (kOutBf16) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        st_shared_u4(srow + (uint32_t)((((h * 4 + q) ^ sw)) * 16),
                                     pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                                     pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        st_shared_u4(srow + (uint32_t)((q ^ sw) * 16), __float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]),
                                     __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3]));
                }
                if (h == kSubsPerChunk - 1) {
                    fence_proxy_async_smem();
                    asm volatile("bar.sync 1, %0;" :: "n"(kEpiThreads) : "memory");
                    if (et == 0) {
                        // pair mode: the upper CTA's 128 rows may lie entirely beyond the last row
                        if (!kPair || m0 < p.M) tma_store_2d(&tmap_c, staging + sbuf * kStagingBytes, n0 + ch * kColsPerChunk, m0);
                        tma_store_commit();
                    }
                    sbuf ^= 1;
                }
            };

            uint32_t va[32], vb[32];
            tmem_ld32(t_row, va);
#pragma unroll 1
            for (int sub = 0; sub < kSubs; sub += 2) {
                tmem_ld_wait(va);
                tmem_ld32(t_row + (uint32_t)((sub + 1) * 32), vb);
                process(va, sub);
                tmem_ld_wait(vb);
                if (sub + 2 < kSubs) tmem_ld32(t_row + (uint32_t)((sub + 2) * 32), va);
                else release_acc();
                process(vb, sub + 1);
            }
            if (n_contrib > 0) {
                // every epilogue thread has read the slots: hand the flags back (0) for the next launch on this workspace
                asm volatile("bar.sync 1, %0;" :: "n"(kEpiThreads) : "memory");
                if (et < n_contrib) ws_flags[slot_of(cta - 1 - et)] = 0u;
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
        if (et == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (kPair) cluster_sync_all();     // no CTA leaves (or frees TMEM) while its peer can still signal or read it
    if (warp == 2) {
        tc_fence_after();
        if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
}

template <bool kOutBf16, int kEpi, bool kPair>
static int launch_sk_one(int grid, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                         const Params& p, float* slots, uint32_t* flags) {
    constexpr int smem = kPair ? gemm_smem_bytes_pair(kPairStages, 1) : gemm_smem_bytes(kStages, 1);
    static_assert(smem <= 227 * 1024, "shared memory exceeds the 227 KB a CTA can opt into");
    auto kern = gemm_streamk_kernel<kOutBf16, kEpi, kPair>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    TASU_CHECK_CUDA(attr_err);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    // Finishing CTAs spin on flags of other CTAs: all must be resident.  grid <= number of SMs with one CTA per SM, so
    // they are whenever the device is not shared; the cooperative attribute makes the launch fail instead of hang
    // otherwise (one-CTA mode; a cluster launch cannot be cooperative, there the grid size is the guarantee).
    cudaLaunchAttribute attr[1];
    if (kPair) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    } else {
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TASU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, p, slots, flags));
    return TASU_OK;
}

template <bool kOutBf16, bool kPair>
static int launch_sk_epi(int epilogue, int grid, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb,
                         const CUtensorMap& mc, const Params& p, float* slots, uint32_t* flags) {
#define TASU_SK(E) launch_sk_one<kOutBf16, E, kPair>(grid, st, ma, mb, mc, p, slots, flags)
    switch (epilogue) {
        case TASU_EPI_NONE: return TASU_SK(TASU_EPI_NONE);
        case TASU_EPI_BIAS: return TASU_SK(TASU_EPI_BIAS);
        case TASU_EPI_BIAS_SILU: return TASU_SK(TASU_EPI_BIAS_SILU);
        case TASU_EPI_BIAS_RELU: return TASU_SK(TASU_EPI_BIAS_RELU);
        case TASU_EPI_LNFOLD_SILU: return TASU_SK(TASU_EPI_LNFOLD_SILU);
        case TASU_EPI_LNFOLD: return TASU_SK(TASU_EPI_LNFOLD);
        default: return TASU_SK(TASU_EPI_SOFTMAX);
    }
#undef TASU_SK
}

}  // namespace gemm
}  // namespace tasu

using namespace tasu;
using namespace tasu::gemm;

extern "C" int64_t tasu_gemm_streamk_workspace(void) {
    const int64_t g = sm_count();
    return g * (int64_t)kSkSlotFloats * 4 + ((g * 4 + 255) / 256) * 256;
}

extern "C" int tasu_gemm_bf16_tn_streamk(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int c_dtype,
                                         int64_t ldc, int M, int N, int K, int epilogue, const float* bias,
                                         const float* row_rstd, const float* row_mean, const float* colsum,
                                         const int32_t* m_dev, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_common(A, lda, B, ldb, C, c_dtype, ldc, M, N, K, epilogue, bias, row_rstd, row_mean, colsum);
    if (rc != TASU_OK || M == 0) return rc;
    const int csz = c_dtype == TASU_F32 ? 4 : 2;
    TASU_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0), "base pointers must be 16-byte aligned");
    TASU_CHECK_ARG((lda * 2) % 16 == 0 && (ldb * 2) % 16 == 0 && (ldc * csz) % 16 == 0, "row pitches must be multiples of 16 bytes");
    TASU_CHECK_ARG(workspace && ((uintptr_t)workspace % 256 == 0) && workspace_bytes >= tasu_gemm_streamk_workspace(),
                   "workspace missing, unaligned or smaller than tasu_gemm_streamk_workspace()");
    // CTA pairs (one 256x256 tile per cluster of two) for problems with more than one 128-row tile, as tasu_gemm_bf16_tn
    const bool pair = option(TASU_OPT_GEMM_PAIR) != 0 && M > BM && sm_count() >= 2;
    CUtensorMap ma, mb, mc;
    rc = make_map(&ma, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, BM, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
    rc = make_map(&mb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldb, pair ? BN / 2 : BN, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
    rc = make_map(&mc, C, c_dtype == TASU_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, csz,
                  M, N, ldc, BM, c_dtype == TASU_F32 ? 32 : 64, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    Params p{M, N, K, m_dev, epilogue, bias, row_rstd, row_mean, colsum};
    // the grid is always the SM count (pair mode: rounded down to whole clusters): with a device-side row count the
    // number of live tiles is not known here, and the flag / slot arrays are indexed by CTA
    const int sms = sm_count();
    const int grid = pair ? (sms / 2) * 2 : sms;
    uint32_t* flags = reinterpret_cast<uint32_t*>(workspace);
    float* slots = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ((sms * 4 + 255) / 256) * 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (pair)
        rc = c_dtype == TASU_BF16 ? launch_sk_epi<true, true>(epilogue, grid, st, ma, mb, mc, p, slots, flags)
                                  : launch_sk_epi<false, true>(epilogue, grid, st, ma, mb, mc, p, slots, flags);
    else
        rc = c_dtype == TASU_BF16 ? launch_sk_epi<true, false>(epilogue, grid, st, ma, mb, mc, p, slots, flags)
                                  : launch_sk_epi<false, false>(epilogue, grid, st, ma, mb, mc, p, slots, flags);
    if (rc) return rc;
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

// HOST: the stream-K schedule of CTA `cta` for a problem of num_tiles output tiles x k_blocks K-blocks on `grid` CTAs
// (the function the kernel runs, exposed so that coverage / ordering properties are tested on the CPU).
// pieces_host: up to 2 x {tile, kb0, kb1, kind, n_contrib}; returns the number of pieces or a negative error.
extern "C" int tasu_gemm_streamk_schedule_host(int num_tiles, int k_blocks, int grid, int cta, int32_t* pieces_host,
                                               int32_t* dp_tiles_host) {
    TASU_CHECK_ARG(num_tiles >= 0 && k_blocks > 0 && grid > 0 && cta >= 0 && cta < grid && pieces_host, "schedule arguments");
    int dp, rem, sk;
    sk_split(num_tiles, k_blocks, grid, &dp, &rem, &sk);
    if (dp_tiles_host) *dp_tiles_host = dp;
    SkPiece out[2];
    const int n = sk_pieces(num_tiles, k_blocks, grid, cta, out);
    for (int i = 0; i < n; ++i) {
        pieces_host[5 * i + 0] = out[i].tile; pieces_host[5 * i + 1] = out[i].kb0; pieces_host[5 * i + 2] = out[i].kb1;
        pieces_host[5 * i + 3] = out[i].kind; pieces_host[5 * i + 4] = out[i].n_contrib;
    }
    return n;
}
