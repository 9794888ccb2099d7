// Step 3: C[M,N] = epilogue(A[M,K] · B[N,K]^T) on the 5th-generation tensor cores (sm_100a).
//
// Replaces cuBLAS addmm at Multitask/model/ps-slm.py:450,581 (ctc_lo, K=512, N=25055) and
// Multitask/model/projector.py:141-143 (Linear(25055,2048)+SiLU with the LayerNorm of :139
// folded into the epilogue, Linear(2048,1536)), :35-37 and :16.
//
// Design (hand-written, no CUTLASS):
//   * persistent: one CTA per SM, static round-robin over 128x256 output tiles, n fastest so
//     the CTAs running at the same time share A tiles and stream the same weight K-slices in L2;
//   * warp 0 = TMA producer (cp.async.bulk.tensor 2D, 128B swizzle, OOB zero-fill handles the
//     ragged M, N=25055 and K=25055 edges), 4-stage mbarrier ring of {A 128x64, B 256x64} bf16;
//   * warp 1 = MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//     (M=128, N=256, K=16) from shared-memory descriptors, accumulating fp32 in TMEM;
//     tcgen05.commit releases smem stages and publishes finished accumulators;
//   * TMEM holds two 128x256 fp32 accumulators (all 512 columns) so the epilogue of tile i
//     overlaps the main loop of tile i+1;
//   * warps 4-7 = epilogue: tcgen05.ld 32 lanes x 32 columns per warp, fused bias / SiLU / ReLU /
//     LayerNorm-fold math in registers, swizzled st.shared into a 128-byte-wide staging tile,
//     TMA store (clips the ragged edges) double-buffered against the next chunk.
#include "gemm_common.cuh"

namespace tasu {
namespace gemm {


// kOutBf16: C is bf16 (64 columns per 128-byte staging row) else fp32 (32 columns)
// kSt smem pipeline stages; kGroups epilogue warp-groups of 4 warps (2 groups interleave the 128-byte column
// chunks of a tile, each with its own staging pair and TMA-store thread: for shallow-K, store-heavy shapes)
// kMajor: bit 0 = A is MN-major ([K, M] in memory), bit 1 = B is MN-major ([K, N] in memory); 0 = both K-major ("TN")
// kPair: CTA-pair mode (launched as clusters of 2): one tcgen05.mma.cta_group::2 of M = 256 per 256x256 pair tile;
//        each CTA stages its own 128 rows of A and 128 of the 256 B rows (32 KB per stage instead of 48 KB, a third
//        less L2->smem traffic per flop), the rank-0 CTA issues the MMAs and multicasts its commits to both CTAs,
//        every CTA runs the epilogue of its own 128 accumulator rows.  EXPERIMENTAL: see tasu_set_option.
// kPrefetch (EXPERIMENTAL, TASU_OPT_EPI_PREFETCH): the epilogue fetches the bias / colsum / row-statistics values of
//        its NEXT tile into registers while it processes the current one, instead of loading them at the top of every
//        tile.  For K = 512 a tile lasts ~2 us and the exposed L2 latency of those loads is the largest single stall
//        of the kept-frame softmax GEMM (profiles/r01h_ncu_detail.md).  Same values, same arithmetic: results are
//        bit-identical to the default kernel.
template <bool kOutBf16, int kEpi, int kSt, int kGroups, int kMajor = 0, bool kPair = false, bool kPrefetch = false>
__global__ void __launch_bounds__(128 + 128 * kGroups, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];          // 128B-swizzle atoms need 1024-byte alignment
    if ((smem_u32(smem) & 1023u) != 0) __trap();               // no static shared memory in this kernel → offset 0
    static_assert(!kPair || kMajor == 0, "pair mode: K-major operands only");
    constexpr int kStages = kSt;
    constexpr int kStgBytes = kPair ? kPairStageBytes : kStageBytes;               // A 16 KB + B 32 KB (pair: 16 KB)
    constexpr int kTileM = kPair ? 2 * BM : BM;                                    // rows of C per work item
    uint8_t* staging = smem + kStages * kStgBytes;                                 // [kGroups][2] x 16 KB
    float* s_aux = reinterpret_cast<float*>(staging + 2 * kGroups * kStagingBytes);  // [kGroups][bias, colsum][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kGroups * kStagingBytes + kGroups * kAuxBytes);
    uint64_t* full_bar = bars;                           // [kStages]
    uint64_t* empty_bar = bars + kStages;                // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;            // [kAccStages]
    uint64_t* tmem_empty = bars + 2 * kStages + kAccStages;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M_live = p.m_dev != nullptr ? min(max(__ldg(p.m_dev), 0), p.M) : p.M;
    const int m_tiles = (M_live + kTileM - 1) / kTileM, n_tiles = (p.N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int k_blocks = (p.K + BK - 1) / BK;
    // work distribution: one CTA (pair mode: one cluster of two CTAs) per tile, static round-robin
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
#define TASU_TILE_LOOP for (int tile = kPair ? (blockIdx.x >> 1) : blockIdx.x; tile < num_tiles; tile += kPair ? (gridDim.x >> 1) : gridDim.x)
#define TASU_TILE_M0 ((tile / n_tiles) * kTileM + (kPair ? (int)rank * BM : 0))

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); prefetch_tmap(&tmap_c);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        // pair mode: the epilogue threads of BOTH CTAs release an accumulator on the MMA-issuing CTA's barrier
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpiThreads * kGroups * (kPair ? 2 : 1)); }
        fence_barrier_init();
    }
    if (warp == 2) {   // whole warp allocates all 512 TMEM columns (1 CTA per SM)
        if (kPair) {   // one warp of each CTA of the pair
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
    if (kPair) cluster_sync_all();     // the peer's barriers are initialised before anything is signalled on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            TASU_TILE_LOOP {
                const int m0 = TASU_TILE_M0, n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kStgBytes;
                    if (kPair) {
                        // both CTAs' halves of the stage complete on the rank-0 barrier the MMA thread waits on; the
                        // peer's bytes may land before rank 0 posts its expect_tx (the pending arrival keeps the phase open)
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
                        const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        tma_load_2d_pair(&tmap_a, fb, sa, kb * BK, m0);
                        tma_load_2d_pair(&tmap_b, fb, sa + kABytes, kb * BK, n0 + (int)rank * (BN / 2));
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(&full_bar[stage], kStageBytes);
                    if (kMajor & 1) {                               // [64 K-rows][64 M] boxes, 8 KB apart
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i) tma_load_2d(&tmap_a, &full_bar[stage], sa + i * 8192, m0 + 64 * i, kb * BK);
                    } else {
                        tma_load_2d(&tmap_a, &full_bar[stage], sa, kb * BK, m0);
                    }
                    if (kMajor & 2) {
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)
                            tma_load_2d(&tmap_b, &full_bar[stage], sa + kABytes + i * 8192, n0 + 64 * i, kb * BK);
                    } else {
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
            constexpr uint32_t idesc = (kPair ? kInstrDescPair : kInstrDesc) | ((uint32_t)(kMajor & 1) << 15) | ((uint32_t)((kMajor >> 1) & 1) << 16);
            // per UMMA_K step: K-major advances 32 B inside the swizzle row (+2 in the >>4 address field); MN-major
            // advances two 8-row K groups = 2048 B (+128)
            constexpr uint64_t a_step = (kMajor & 1) ? 128 : 2, b_step = (kMajor & 2) ? 128 : 2;
            TASU_TILE_LOOP {
                if (kPair) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);   // drained by both CTAs' epilogues
                else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStgBytes);
                    const uint64_t adesc = (kMajor & 1) ? make_smem_desc_mn(sa) : make_smem_desc(sa);
                    const uint64_t bdesc = (kMajor & 2) ? make_smem_desc_mn(sa + kABytes) : make_smem_desc(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (kPair) umma_bf16_pair(d_tmem, adesc + a_step * (uint64_t)k, bdesc + b_step * (uint64_t)k, idesc,
                                                  (kb > 0 || k > 0) ? 1u : 0u);
                        else umma_bf16(d_tmem, adesc + a_step * (uint64_t)k, bdesc + b_step * (uint64_t)k, idesc,
                                       (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    if (kPair) {                                    // the same barriers of both CTAs
                        umma_commit_pair(&empty_bar[stage]);
                        if (kb == k_blocks - 1) umma_commit_pair(&tmem_full[acc]);
                    } else {
                        umma_commit(&empty_bar[stage]);             // smem stage reusable once these MMAs retire
                        if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int grp = (warp - 4) >> 2;                           // epilogue group: chunks grp, grp+kGroups, ...
        const int ew = (warp - 4) & 3;                             // TMEM lanes [32*ew, 32*ew+32)
        const int et = ew * 32 + lane;                             // 0..127 = row of the tile
        float* s_bias = s_aux + grp * (2 * BN);
        float* s_colsum = s_bias + BN;
        uint8_t* gstaging = staging + grp * 2 * kStagingBytes;
        const int bar_id = 1 + grp;
        constexpr int kSubs = BN / 32;                             // 32-column TMEM loads per tile
        constexpr int kSubsPerChunk = kOutBf16 ? 2 : 1;            // one staging chunk = 128 bytes of C per row
        constexpr int kColsPerChunk = 32 * kSubsPerChunk;
        const uint32_t stg_u32 = smem_u32(gstaging);
        const int sw = et & 7;
        int acc = 0; uint32_t acc_phase = 0;
        int sbuf = 0;
        // kPrefetch: values of the NEXT tile, fetched one tile ahead (column c = et + i * kEpiThreads of the tile)
        constexpr bool kLnFold = kEpi == TASU_EPI_LNFOLD_SILU || kEpi == TASU_EPI_LNFOLD;
        constexpr bool kRowVec = kLnFold || kEpi == TASU_EPI_SOFTMAX;
        float pf_bias[BN / kEpiThreads], pf_colsum[BN / kEpiThreads], pf_rstd = 1.f, pf_mean = 0.f;
        auto prefetch_tile = [&](int t) {
            const int pm0 = (t / n_tiles) * kTileM + (kPair ? (int)rank * BM : 0), pn0 = (t % n_tiles) * BN;
#pragma unroll
            for (int i = 0; i < BN / kEpiThreads; ++i) {
                const int col = pn0 + et + i * kEpiThreads;
                pf_bias[i] = (kEpi != TASU_EPI_NONE && col < p.N) ? __ldg(p.bias + col) : 0.f;
                pf_colsum[i] = (kLnFold && col < p.N) ? __ldg(p.colsum + col) : 0.f;
            }
            pf_rstd = 1.f; pf_mean = 0.f;
            if (kRowVec && pm0 + et < M_live) { pf_rstd = __ldg(p.row_rstd + pm0 + et); pf_mean = __ldg(p.row_mean + pm0 + et); }
        };
        if (kPrefetch) {
            const int t0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
            if (t0 < num_tiles) prefetch_tile(t0);
        }
        TASU_TILE_LOOP {
            const int m0 = TASU_TILE_M0, n0 = (tile % n_tiles) * BN;
            const int grow = m0 + et;
            // per-column vectors of this tile → shared memory (read back as broadcast LDS.128).
            // Safe to overwrite: every read of the previous tile happened before its last bar.sync.
            if (kPrefetch) {
                if (kEpi != TASU_EPI_NONE) {
#pragma unroll
                    for (int i = 0; i < BN / kEpiThreads; ++i) {
                        s_bias[et + i * kEpiThreads] = pf_bias[i] * (kEpi == TASU_EPI_SOFTMAX ? kLog2e : 1.f);
                        if (kLnFold) s_colsum[et + i * kEpiThreads] = pf_colsum[i];
                    }
                }
            } else
            if (kEpi != TASU_EPI_NONE) {
                for (int c = et; c < BN; c += kEpiThreads) {   // each group stages its own copy (own barrier)
                    const int col = n0 + c;
                    // softmax epilogue works in the log2 domain: bias pre-scaled once per tile, not once per element
                    s_bias[c] = col < p.N ? __ldg(p.bias + col) * (kEpi == TASU_EPI_SOFTMAX ? kLog2e : 1.f) : 0.f;
                    if (kEpi == TASU_EPI_LNFOLD_SILU || kEpi == TASU_EPI_LNFOLD) s_colsum[c] = col < p.N ? __ldg(p.colsum + col) : 0.f;
                }
            }
            float rstd = 1.f, nmean = 0.f;
            if (kPrefetch) {
                rstd = pf_rstd; nmean = -pf_mean;
                // the next tile's values are in flight while this tile is processed
                const int nxt = tile + (kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x);
                if (nxt < num_tiles) prefetch_tile(nxt);
            } else
            if ((kEpi == TASU_EPI_LNFOLD_SILU || kEpi == TASU_EPI_LNFOLD || kEpi == TASU_EPI_SOFTMAX) && grow < M_live) {
                rstd = __ldg(p.row_rstd + grow); nmean = -__ldg(p.row_mean + grow);
            }
            // softmax: p = exp(acc + bias - max) / sum = 2^(acc*log2e + bias*log2e + rowc), rowc = -max*log2e + log2(1/sum):
            // one FADD + one FFMA + one MUFU.EX2 per element
            const float rowc = kEpi == TASU_EPI_SOFTMAX ? fmaf(nmean, kLog2e, __log2f(fmaxf(rstd, 1e-37f))) : 0.f;
            const int n_chunks = min(BN / kColsPerChunk, (p.N - n0 + kColsPerChunk - 1) / kColsPerChunk);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);

            // one 32-column slab: epilogue math in registers, swizzled st.shared, TMA store per chunk
            auto process = [&](uint32_t (&v)[32], int sub) {
                const int ch = sub / kSubsPerChunk, h = sub % kSubsPerChunk;
                if (ch >= n_chunks) return;                            // fully clipped (uniform over the group's 4 warps)
                const uint32_t srow = stg_u32 + (uint32_t)(sbuf * kStagingBytes + et * 128);
                if (h == 0) {
                    // the TMA store that last read this staging buffer must have finished reading it
                    if (et == 0) tma_store_wait_read<1>();
                    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "n"(kEpiThreads) : "memory");
                }
                float f[32];
                const float4* b4 = reinterpret_cast<const float4*>(s_bias + sub * 32);
                const float4* c4 = reinterpret_cast<const float4*>(s_colsum + sub * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float x[4] = {__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                  __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])};
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
                            for (int e = 0; e < 4; ++e)      // softmax with known row max (-nmean) and 1/sum (rstd)
                                x[e] = ex2_approx(fmaf(x[e], kLog2e, b[e] + rowc));
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
                    for (int q = 0; q < 4; ++q)        // 32 columns → 64 bytes → 16-byte pieces 4h .. 4h+3
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
                    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "n"(kEpiThreads) : "memory");
                    if (et == 0) {
                        // pair mode: the upper CTA's 128 rows may lie entirely beyond the last row
                        if (!kPair || m0 < p.M) tma_store_2d(&tmap_c, gstaging + sbuf * kStagingBytes, n0 + ch * kColsPerChunk, m0);
                        tma_store_commit();
                    }
                    sbuf ^= 1;
                }
            };

            // slabs of this group: chunks grp, grp + kGroups, ... (kSubsPerChunk slabs each), software-pipelined
            constexpr int kGroupSubs = kSubs / kGroups;
            auto sub_of = [&](int it) { return ((it / kSubsPerChunk) * kGroups + grp) * kSubsPerChunk + it % kSubsPerChunk; };
            uint32_t va[32], vb[32];
            tmem_ld32(t_row + (uint32_t)(sub_of(0) * 32), va);
#pragma unroll 1
            for (int it = 0; it < kGroupSubs; it += 2) {
                tmem_ld_wait(va);
                tmem_ld32(t_row + (uint32_t)(sub_of(it + 1) * 32), vb);   // in flight while `va` is processed
                process(va, sub_of(it));
                tmem_ld_wait(vb);
                if (it + 2 < kGroupSubs) tmem_ld32(t_row + (uint32_t)(sub_of(it + 2) * 32), va);
                else {
                    // every tcgen05.ld of this thread for this accumulator has completed → hand it back early
                    tc_fence_before();
                    if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
                    else mbar_arrive(&tmem_empty[acc]);
                }
                process(vb, sub_of(it + 1));
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
#undef TASU_TILE_LOOP
#undef TASU_TILE_M0
}

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

constexpr int kStatsPairStages = 6;                  // pair mode: 6 x 16 KB half-B stages next to the resident A tile
constexpr int kStatsSmemBytesPair = 8 * kABytes + kStatsPairStages * kPairBBytes + 2 * BN * 4 + 256;

// kARes (K <= 512): the 128-frame A tile stays resident in shared memory for the whole vocabulary sweep of an item
// and only the weight tiles stream through a 3-stage ring — a third less L2→SM traffic per MMA.
// kPair (EXPERIMENTAL, with kARes; launched as clusters of 2, see gemm_bf16_tn_kernel): a work item covers 256 frames,
// each CTA keeps its own 128 frames resident and streams HALF of every 256-column weight tile (16 KB per stage, 6
// stages), the rank-0 CTA issues tcgen05.mma.cta_group::2, every CTA reduces the statistics of its own 128 frames.
// kPrefetch (EXPERIMENTAL, TASU_OPT_EPI_PREFETCH bit 1): every epilogue thread fetches its bias value of the NEXT
// vocabulary tile while the current one is reduced (the load at the top of each tile is exposed L2 latency otherwise).
template <bool kARes, bool kPair = false, bool kPrefetch = false>
__global__ void __launch_bounds__(kStatsThreads, 1)
ctc_stats_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const StatsParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    static_assert(!kPair || kARes, "pair mode keeps the A tile resident");
    constexpr int kSt = kPair ? kStatsPairStages : kARes ? 3 : kStages;        // ring stages
    constexpr int kRingStage = kPair ? kPairBBytes : kARes ? kBBytes : kStageBytes;   // bytes per ring stage
    constexpr int kRingOff = kARes ? 8 * kABytes : 0;                          // resident A: 8 k-blocks x 16 KB
    constexpr int kTileM = kPair ? 2 * BM : BM;                                // frames per work item
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
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
#define TASU_ITEM_LOOP for (int item = kPair ? (blockIdx.x >> 1) : blockIdx.x; item < num_items; item += kPair ? (gridDim.x >> 1) : gridDim.x)
#define TASU_ITEM_M0 ((item / p.splits) * kTileM + (kPair ? (int)rank * BM : 0))

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kSt; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kStatsEpiThreads * (kPair ? 2 : 1)); }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
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
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, a_phase = 0;
            TASU_ITEM_LOOP {
                const int m0 = TASU_ITEM_M0;
                const int nb = (item % p.splits) * p.nt_per, ne = min(nb + p.nt_per, n_tiles);
                if (kPair) {
                    // both CTAs' frames / weight halves complete on the rank-0 barriers the MMA thread waits on
                    mbar_wait(a_empty, a_phase ^ 1);
                    if (rank == 0) mbar_expect_tx(a_full, (uint32_t)(2 * k_blocks * kABytes));
                    const uint32_t af = mapa_u32(smem_u32(a_full), 0);
                    for (int kb = 0; kb < k_blocks; ++kb) tma_load_2d_pair(&tmap_a, af, smem + kb * kABytes, kb * BK, m0);
                    a_phase ^= 1;
                    for (int nt = nb; nt < ne; ++nt) {
                        for (int kb = 0; kb < k_blocks; ++kb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            if (rank == 0) mbar_expect_tx(&full_bar[stage], (uint32_t)(2 * kRingStage));
                            tma_load_2d_pair(&tmap_b, mapa_u32(smem_u32(&full_bar[stage]), 0), ring + stage * kRingStage,
                                             kb * BK, nt * BN + (int)rank * (BN / 2));
                            if (++stage == kSt) { stage = 0; phase ^= 1; }
                        }
                    }
                    continue;
                }
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
        if (lane == 0 && (!kPair || rank == 0)) {
            int stage = 0; uint32_t phase = 0, a_phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            TASU_ITEM_LOOP {
                const int nb = (item % p.splits) * p.nt_per, ne = min(nb + p.nt_per, n_tiles);
                if (kARes) { mbar_wait(a_full, a_phase); tc_fence_after(); a_phase ^= 1; }
                for (int nt = nb; nt < ne; ++nt) {
                    if (kPair) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
                    else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
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
                            if (kPair) umma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDescPair,
                                                      (kb > 0 || k > 0) ? 1u : 0u);
                            else umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDesc,
                                           (kb > 0 || k > 0) ? 1u : 0u);
                        }
                        if (kPair) {
                            umma_commit_pair(&empty_bar[stage]);
                            if (kb == k_blocks - 1) umma_commit_pair(&tmem_full[acc]);
                        } else {
                            umma_commit(&empty_bar[stage]);
                            if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
                        }
                        if (++stage == kSt) { stage = 0; phase ^= 1; }
                    }
                    if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
                }
                if (kPair) umma_commit_pair(a_empty);              // both CTAs may replace their resident frames
                else if (kARes) umma_commit(a_empty);              // every MMA of this item has retired → A may be replaced
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
        // kPrefetch: this thread's bias value (column et) of the tile after the current one, in program order
#define TASU_BIAS_OF(n0_) (((n0_) + et) < p.N ? (p.bias ? __ldg(p.bias + (n0_) + et) : 0.f) : -INFINITY)
        [[maybe_unused]] float pf_bias = 0.f;
        if constexpr (kPrefetch) {
            const int it0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
            if (it0 < num_items) pf_bias = TASU_BIAS_OF((it0 % p.splits) * p.nt_per * BN);
        }
        TASU_ITEM_LOOP {
            const int split = item % p.splits;
            const int row = TASU_ITEM_M0 + ew * 32 + lane;
            const int nb = split * p.nt_per, ne = min(nb + p.nt_per, n_tiles);
            float rm = -INFINITY, rs = 0.f, rs2 = 0.f, xb = 0.f;
            int best = 0x7fffffff;
            for (int nt = nb; nt < ne; ++nt) {
                const int n0 = nt * BN;
                float* sb = s_bias + bbuf * BN;
                if constexpr (kPrefetch) {
                    sb[et] = pf_bias;
                    // next tile of this item, else the first tile of this CTA's next item
                    const int nxt_item = item + (kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x);
                    if (nt + 1 < ne) pf_bias = TASU_BIAS_OF((nt + 1) * BN);
                    else if (nxt_item < num_items) pf_bias = TASU_BIAS_OF((nxt_item % p.splits) * p.nt_per * BN);
                } else
                {
                    const int col = n0 + et;                      // -inf masks the columns beyond the vocabulary
                    sb[et] = col < p.N ? (p.bias ? __ldg(p.bias + col) : 0.f) : -INFINITY;
                }
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
                        if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
                        else mbar_arrive(&tmem_empty[acc]);
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
    if (kPair) cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
#undef TASU_ITEM_LOOP
#undef TASU_ITEM_M0
#undef TASU_BIAS_OF
}

// ------------------------------------------------------------------ EXPERIMENTAL: 16 epilogue warps
// TASU_OPT_STATS_WIDE.  profiles/r01h_ncu_detail.md: ctc_stats_kernel is epilogue-bound (the MMA thread spins ~4 M times
// on tmem_empty) although MUFU (42 %) and the issue slots (46 %) are half idle — with two epilogue warps per scheduler the
// fixed-latency dependency waits are exposed.  This variant runs FOUR epilogue warps per scheduler: 16 warps, each owning
// one TMEM lane quadrant and a 64-column quarter of the 256-column tile, read as 16-column tcgen05.ld slabs so that
// 640 threads fit the register file; the bias value of the next tile is prefetched.  Producer / MMA roles, the resident
// A tile and the 3-stage weight ring are those of ctc_stats_kernel<true>; partials are written per (split, quarter).
constexpr int kWideThreads = 640;                    // warps 0-3 control, warps 4-19 epilogue
constexpr int kWideEpiThreads = 512;

__global__ void __launch_bounds__(kWideThreads, 1)
ctc_stats_wide_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const StatsParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    constexpr int kSt = 3;
    uint8_t* ring = smem + 8 * kABytes;                                        // resident A: 8 k-blocks x 16 KB
    float* s_bias = reinterpret_cast<float*>(ring + kSt * kBBytes);            // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kSt * kBBytes + 2 * BN * 4);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kSt;
    uint64_t* tmem_full = bars + 2 * kSt;
    uint64_t* tmem_empty = bars + 2 * kSt + kAccStages;
    uint64_t* a_full = bars + 2 * kSt + 2 * kAccStages;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(a_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
    const int num_items = m_tiles * p.splits;
    const int k_blocks = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kSt; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kWideEpiThreads); }
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
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int m0 = (item / p.splits) * BM;
                const int nb = (item % p.splits) * p.nt_per, ne = min(nb + p.nt_per, n_tiles);
                mbar_wait(a_empty, a_phase ^ 1);                   // MMAs of the previous item are done with A
                mbar_expect_tx(a_full, (uint32_t)(k_blocks * kABytes));
                for (int kb = 0; kb < k_blocks; ++kb) tma_load_2d(&tmap_a, a_full, smem + kb * kABytes, kb * BK, m0);
                a_phase ^= 1;
                for (int nt = nb; nt < ne; ++nt) {
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], (uint32_t)kBBytes);
                        tma_load_2d(&tmap_b, &full_bar[stage], ring + stage * kBBytes, kb * BK, nt * BN);
                        if (++stage == kSt) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, a_phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int nb = (item % p.splits) * p.nt_per, ne = min(nb + p.nt_per, n_tiles);
                mbar_wait(a_full, a_phase); tc_fence_after(); a_phase ^= 1;
                for (int nt = nb; nt < ne; ++nt) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint64_t adesc = make_smem_desc(smem_u32(smem + kb * kABytes));
                        const uint64_t bdesc = make_smem_desc(smem_u32(ring + stage * kBBytes));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDesc,
                                      (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&empty_bar[stage]);
                        if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
                        if (++stage == kSt) { stage = 0; phase ^= 1; }
                    }
                    if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
                }
                umma_commit(a_empty);
            }
        }
    } else if (warp >= 4) {
        // 16 epilogue warps: warps (q, q+4, q+8, q+12) share TMEM lane quadrant q; `quarter` selects 64 of the 256 columns
        const int ew = (warp - 4) & 3, quarter = (warp - 4) >> 2;
        const int et = threadIdx.x - 128;                          // 0..511; threads 0..255 stage the bias
        constexpr float kL2e = 1.4426950408889634f;
        int acc = 0; uint32_t acc_phase = 0;
        int bbuf = 0;
        auto bias_of = [&](int n0) {                               // -inf masks the columns beyond the vocabulary
            const int col = n0 + et;
            return col < p.N ? (p.bias ? __ldg(p.bias + col) : 0.f) : -INFINITY;
        };
        float pf_bias = 0.f;                                       // bias of the NEXT tile (column et), fetched one tile ahead
        if (et < BN && (int)blockIdx.x < num_items) pf_bias = bias_of(((int)blockIdx.x % p.splits) * p.nt_per * BN);
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int m_tile = item / p.splits, split = item % p.splits;
            const int row = m_tile * BM + ew * 32 + lane;
            const int nb = split * p.nt_per, ne = min(nb + p.nt_per, n_tiles);
            float rm = -INFINITY, rs = 0.f, rs2 = 0.f, xb = 0.f;
            int best = 0x7fffffff;
            for (int nt = nb; nt < ne; ++nt) {
                const int n0 = nt * BN;
                float* sb = s_bias + bbuf * BN;
                if (et < BN) {
                    sb[et] = pf_bias;
                    const int nxt_item = item + (int)gridDim.x;
                    if (nt + 1 < ne) pf_bias = bias_of((nt + 1) * BN);
                    else if (nxt_item < num_items) pf_bias = bias_of((nxt_item % p.splits) * p.nt_per * BN);
                }
                asm volatile("bar.sync 1, %0;" :: "n"(kWideEpiThreads) : "memory");
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);

                auto process = [&](uint32_t (&v)[16], int sub) {   // sub = index of the 16-column slab in the tile
                    const int base = n0 + sub * 16;
                    if (base >= p.N) return;
                    const float4* b4 = reinterpret_cast<const float4*>(sb + sub * 16);
                    float x[16];
                    float cm = -INFINITY;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bb = b4[q];
                        x[4 * q] = __uint_as_float(v[4 * q]) + bb.x;
                        x[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + bb.y;
                        x[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + bb.z;
                        x[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + bb.w;
                        cm = fmaxf(cm, fmaxf(fmaxf(x[4 * q], x[4 * q + 1]), fmaxf(x[4 * q + 2], x[4 * q + 3])));
                    }
                    if (cm > rm) {
                        const float f = ex2_approx((rm - cm) * kL2e);      // rm = -inf on the first slab: 2^-inf = 0
                        rs *= f;
                        rs2 *= f * f;
                        rm = cm;
                        int j0 = 15;
#pragma unroll
                        for (int j = 14; j >= 0; --j) j0 = (x[j] == cm) ? j : j0;
                        best = base + j0;                          // first index of the maximum (torch tie rule)
                    }
                    const float mb = rm * kL2e;
                    float acc_s = 0.f, acc_s2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float e = ex2_approx(fmaf(x[j], kL2e, -mb));
                        acc_s += e;
                        acc_s2 = fmaf(e, e, acc_s2);
                    }
                    rs += acc_s;
                    rs2 += acc_s2;
                    if (p.blank >= base && p.blank < base + 16) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) xb = (base + j == p.blank) ? x[j] : xb;
                    }
                };

                uint32_t va[16], vb[16];
                const int sub0 = quarter * 4, sub1 = sub0 + 4;             // this warp's four 16-column slabs
                tmem_ld16(t_row + (uint32_t)(sub0 * 16), va);
#pragma unroll 1
                for (int sub = sub0; sub < sub1; sub += 2) {
                    tmem_ld_wait16(va);
                    tmem_ld16(t_row + (uint32_t)((sub + 1) * 16), vb);
                    process(va, sub);
                    tmem_ld_wait16(vb);
                    if (sub + 2 < sub1) tmem_ld16(t_row + (uint32_t)((sub + 2) * 16), va);
                    else { tc_fence_before(); mbar_arrive(&tmem_empty[acc]); }
                    process(vb, sub + 1);
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
                bbuf ^= 1;
            }
            if (row < p.M) {
                const int64_t o = (int64_t)(split * 4 + quarter) * p.M + row;  // ascending column order of the partials
                p.part_max[o] = rm;
                p.part_sum[o] = rs;
                p.part_sum2[o] = rs2;
                p.part_arg[o] = best;
                const int bt = p.blank / BN, bq = (p.blank % BN) / (BN / 4);
                if (bt >= nb && bt < ne && bq == quarter) p.xb_raw[row] = xb;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
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
    float m = -INFINITY; int a = 0x7fffffff;
    const int parts = 2 * p.splits;                                        // (split, column half), ascending columns
    for (int s = 0; s < parts; ++s) {
        const float pm = p.part_max[(int64_t)s * p.M + r];
        if (pm > m) { m = pm; a = p.part_arg[(int64_t)s * p.M + r]; }      // ties keep the lower part = lower index
    }
    float sum = 0.f, sum2 = 0.f;
    for (int s = 0; s < parts; ++s) {
        const float f = exp2f((p.part_max[(int64_t)s * p.M + r] - m) * 1.4426950408889634f);
        sum += p.part_sum[(int64_t)s * p.M + r] * f;
        sum2 += p.part_sum2[(int64_t)s * p.M + r] * f * f;
    }
    if (row_sumexp2) row_sumexp2[i] = sum2;
    argmax[i] = a;
    x_blank[i] = p.xb_raw[r];
    row_max[i] = m;
    row_sumexp[i] = sum;
}

// the same merge for ctc_stats_wide_kernel: four column quarters per split
__global__ void __launch_bounds__(256)
ctc_stats_combine_wide_kernel(StatsParams p, int B, int T, int n_prefix, int32_t* __restrict__ argmax,
                              float* __restrict__ x_blank, float* __restrict__ row_max, float* __restrict__ row_sumexp,
                              float* __restrict__ row_sumexp2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * T) return;
    const int b = (int)(i / T), t = (int)(i % T);
    const int64_t r = (int64_t)b * (T + n_prefix) + n_prefix + t;
    float m = -INFINITY; int a = 0x7fffffff;
    const int parts = 4 * p.splits;                                        // (split, column quarter), ascending columns
    for (int s = 0; s < parts; ++s) {
        const float pm = p.part_max[(int64_t)s * p.M + r];
        if (pm > m) { m = pm; a = p.part_arg[(int64_t)s * p.M + r]; }      // ties keep the lower part = lower index
    }
    float sum = 0.f, sum2 = 0.f;
    for (int s = 0; s < parts; ++s) {
        const float f = exp2f((p.part_max[(int64_t)s * p.M + r] - m) * 1.4426950408889634f);
        sum += p.part_sum[(int64_t)s * p.M + r] * f;
        sum2 += p.part_sum2[(int64_t)s * p.M + r] * f * f;
    }
    if (row_sumexp2) row_sumexp2[i] = sum2;
    argmax[i] = a;
    x_blank[i] = p.xb_raw[r];
    row_max[i] = m;
    row_sumexp[i] = sum;
}

// ------------------------------------------------------------------ CUDA-core cross-check
template <typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, int64_t lda, const __nv_bfloat16* __restrict__ B, int64_t ldb,
                 TC* __restrict__ C, int64_t ldc, Params p) {
    __shared__ float sa[16][17], sb[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
    float acc = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += 16) {
        const int am = blockIdx.y * 16 + ty, bn = blockIdx.x * 16 + ty, kk = k0 + tx;
        sa[ty][tx] = (am < p.M && kk < p.K) ? __bfloat162float(A[(int64_t)am * lda + kk]) : 0.f;
        sb[ty][tx] = (bn < p.N && kk < p.K) ? __bfloat162float(B[(int64_t)bn * ldb + kk]) : 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sb[tx][k], acc);
        __syncthreads();
    }
    if (m < p.M && n < p.N) {
        float x = acc;
        if (p.epilogue == TASU_EPI_LNFOLD_SILU) x = silu_f(fmaf(p.row_rstd[m], x - p.row_mean[m] * p.colsum[n], p.bias[n]));
        else if (p.epilogue == TASU_EPI_LNFOLD) x = fmaf(p.row_rstd[m], x - p.row_mean[m] * p.colsum[n], p.bias[n]);
        else if (p.epilogue == TASU_EPI_SOFTMAX) x = __expf(x + p.bias[n] - p.row_mean[m]) * p.row_rstd[m];
        else if (p.epilogue != TASU_EPI_NONE && p.epilogue != TASU_EPI_SOFTMAX) {
            x += p.bias[n];
            if (p.epilogue == TASU_EPI_BIAS_SILU) x = silu_f(x);
            else if (p.epilogue == TASU_EPI_BIAS_RELU) x = fmaxf(x, 0.f);
        }
        C[(int64_t)m * ldc + n] = from_f32<TC>(x);
    }
}



template <bool kOutBf16, int kEpi, int kSt, int kGroups>
static int launch_one(int grid, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                      const Params& p) {
    constexpr int smem = gemm_smem_bytes(kSt, kGroups);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(gemm_bf16_tn_kernel<kOutBf16, kEpi, kSt, kGroups>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    });
    TASU_CHECK_CUDA(attr_err);
    gemm_bf16_tn_kernel<kOutBf16, kEpi, kSt, kGroups><<<grid, 128 + 128 * kGroups, smem, st>>>(ma, mb, mc, p);
    return TASU_OK;
}

// EXPERIMENTAL (TASU_OPT_EPI_PREFETCH bit 0): shallow-K configuration with the epilogue vectors fetched one tile ahead
template <bool kOutBf16, int kEpi>
static int launch_one_prefetch(int grid, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                               const Params& p) {
    constexpr int smem = gemm_smem_bytes(3, 2);
    auto kern = gemm_bf16_tn_kernel<kOutBf16, kEpi, 3, 2, 0, false, true>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    TASU_CHECK_CUDA(attr_err);
    kern<<<grid, 128 + 128 * 2, smem, st>>>(ma, mb, mc, p);
    return TASU_OK;
}

// deep-K shapes: 4 smem stages, one epilogue group; shallow-K (store-heavy) shapes: 3 stages, two epilogue groups
template <bool kOutBf16, int kEpi>
static int launch_cfg(bool shallow_k, int grid, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb,
                      const CUtensorMap& mc, const Params& p) {
    if (shallow_k && (option(TASU_OPT_EPI_PREFETCH) & 1) != 0) return launch_one_prefetch<kOutBf16, kEpi>(grid, st, ma, mb, mc, p);
    return shallow_k ? launch_one<kOutBf16, kEpi, 3, 2>(grid, st, ma, mb, mc, p)
                     : launch_one<kOutBf16, kEpi, 4, 1>(grid, st, ma, mb, mc, p);
}

template <bool kOutBf16>
static int launch_epi(int epilogue, bool sk, int grid, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb,
                      const CUtensorMap& mc, const Params& p) {
    switch (epilogue) {
        case TASU_EPI_NONE: return launch_cfg<kOutBf16, TASU_EPI_NONE>(sk, grid, st, ma, mb, mc, p);
        case TASU_EPI_BIAS: return launch_cfg<kOutBf16, TASU_EPI_BIAS>(sk, grid, st, ma, mb, mc, p);
        case TASU_EPI_BIAS_SILU: return launch_cfg<kOutBf16, TASU_EPI_BIAS_SILU>(sk, grid, st, ma, mb, mc, p);
        case TASU_EPI_BIAS_RELU: return launch_cfg<kOutBf16, TASU_EPI_BIAS_RELU>(sk, grid, st, ma, mb, mc, p);
        case TASU_EPI_LNFOLD_SILU: return launch_cfg<kOutBf16, TASU_EPI_LNFOLD_SILU>(sk, grid, st, ma, mb, mc, p);
        case TASU_EPI_LNFOLD: return launch_cfg<kOutBf16, TASU_EPI_LNFOLD>(sk, grid, st, ma, mb, mc, p);
        default: return launch_cfg<kOutBf16, TASU_EPI_SOFTMAX>(sk, grid, st, ma, mb, mc, p);
    }
}

static int launch_dispatch(bool out_bf16, int epilogue, bool shallow_k, int grid, cudaStream_t st, const CUtensorMap& ma,
                           const CUtensorMap& mb, const CUtensorMap& mc, const Params& p) {
    return out_bf16 ? launch_epi<true>(epilogue, shallow_k, grid, st, ma, mb, mc, p)
                    : launch_epi<false>(epilogue, shallow_k, grid, st, ma, mb, mc, p);
}

// CTA-pair mode (EXPERIMENTAL, TASU_OPT_GEMM_PAIR): clusters of two CTAs, one 256x256 tile per cluster
template <bool kOutBf16, int kEpi, int kSt, int kGroups>
static int launch_pair_one(int clusters, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb,
                           const CUtensorMap& mc, const Params& p) {
    constexpr int smem = gemm_smem_bytes_pair(kSt, kGroups);
    static_assert(smem <= 227 * 1024, "pair-mode shared memory exceeds the 227 KB a CTA can opt into");
    static_assert((2 * kSt + 2 * kAccStages) * 8 + 4 <= 256, "barrier area");
    auto kern = gemm_bf16_tn_kernel<kOutBf16, kEpi, kSt, kGroups, 0, true>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    TASU_CHECK_CUDA(attr_err);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2u * (unsigned)clusters);
    cfg.blockDim = dim3(128 + 128 * kGroups);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TASU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, p));
    return TASU_OK;
}

template <bool kOutBf16, int kEpi>
static int launch_pair_cfg(bool shallow_k, int clusters, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb,
                           const CUtensorMap& mc, const Params& p) {
    return shallow_k ? launch_pair_one<kOutBf16, kEpi, kPairStagesShallow, 2>(clusters, st, ma, mb, mc, p)
                     : launch_pair_one<kOutBf16, kEpi, kPairStages, 1>(clusters, st, ma, mb, mc, p);
}

template <bool kOutBf16>
static int launch_pair_epi(int epilogue, bool sk, int clusters, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb,
                           const CUtensorMap& mc, const Params& p) {
    switch (epilogue) {
        case TASU_EPI_NONE: return launch_pair_cfg<kOutBf16, TASU_EPI_NONE>(sk, clusters, st, ma, mb, mc, p);
        case TASU_EPI_BIAS: return launch_pair_cfg<kOutBf16, TASU_EPI_BIAS>(sk, clusters, st, ma, mb, mc, p);
        case TASU_EPI_BIAS_SILU: return launch_pair_cfg<kOutBf16, TASU_EPI_BIAS_SILU>(sk, clusters, st, ma, mb, mc, p);
        case TASU_EPI_BIAS_RELU: return launch_pair_cfg<kOutBf16, TASU_EPI_BIAS_RELU>(sk, clusters, st, ma, mb, mc, p);
        case TASU_EPI_LNFOLD_SILU: return launch_pair_cfg<kOutBf16, TASU_EPI_LNFOLD_SILU>(sk, clusters, st, ma, mb, mc, p);
        case TASU_EPI_LNFOLD: return launch_pair_cfg<kOutBf16, TASU_EPI_LNFOLD>(sk, clusters, st, ma, mb, mc, p);
        default: return launch_pair_cfg<kOutBf16, TASU_EPI_SOFTMAX>(sk, clusters, st, ma, mb, mc, p);
    }
}

template <int kSt, int kGroups, int kMajor>
static int launch_major(int grid, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                        const Params& p) {
    constexpr int smem = gemm_smem_bytes(kSt, kGroups);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(gemm_bf16_tn_kernel<false, TASU_EPI_NONE, kSt, kGroups, kMajor>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    });
    TASU_CHECK_CUDA(attr_err);
    gemm_bf16_tn_kernel<false, TASU_EPI_NONE, kSt, kGroups, kMajor><<<grid, 128 + 128 * kGroups, smem, st>>>(ma, mb, mc, p);
    return TASU_OK;
}

}  // namespace gemm
}  // namespace tasu

using namespace tasu;
using namespace tasu::gemm;

extern "C" int tasu_gemm_bf16_f32(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                                  float* C, int64_t ldc, int M, int N, int K, void* stream) {
    TASU_CHECK_ARG(M >= 0 && N > 0 && K > 0, "M >= 0, N,K > 0");
    TASU_CHECK_ARG((a_mn_major == 0 || a_mn_major == 1) && (b_mn_major == 0 || b_mn_major == 1), "major flags are 0 or 1");
    TASU_CHECK_ARG(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K) && ldc >= N, "leading dimension too small");
    if (M == 0) return TASU_OK;
    if (!a_mn_major && !b_mn_major)
        return tasu_gemm_bf16_tn(A, lda, B, ldb, C, TASU_F32, ldc, M, N, K, TASU_EPI_NONE, nullptr, nullptr, nullptr, nullptr,
                                 nullptr, stream);
    TASU_CHECK_ARG(A && B && C, "null pointer");
    TASU_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0), "base pointers must be 16-byte aligned");
    TASU_CHECK_ARG((lda * 2) % 16 == 0 && (ldb * 2) % 16 == 0 && (ldc * 4) % 16 == 0, "row pitches must be multiples of 16 bytes");
    CUtensorMap ma, mb, mc;
    int rc = a_mn_major ? make_map(&ma, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, M, lda, BK, 64, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)
                        : make_map(&ma, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, BM, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
    rc = b_mn_major ? make_map(&mb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, N, ldb, BK, 64, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)
                    : make_map(&mb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldb, BN, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
    rc = make_map(&mc, C, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldc, BM, 32, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    Params p{M, N, K, nullptr, TASU_EPI_NONE, nullptr, nullptr, nullptr, nullptr};
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int grid = sm_count();
    if (grid > tiles) grid = tiles;
    cudaStream_t st = (cudaStream_t)stream;
    const bool sk = K <= 1024;
    const int major = a_mn_major | (b_mn_major << 1);
    if (major == 1) rc = sk ? launch_major<3, 2, 1>(grid, st, ma, mb, mc, p) : launch_major<4, 1, 1>(grid, st, ma, mb, mc, p);
    else if (major == 2) rc = sk ? launch_major<3, 2, 2>(grid, st, ma, mb, mc, p) : launch_major<4, 1, 2>(grid, st, ma, mb, mc, p);
    else rc = sk ? launch_major<3, 2, 3>(grid, st, ma, mb, mc, p) : launch_major<4, 1, 3>(grid, st, ma, mb, mc, p);
    if (rc) return rc;
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_gemm_bf16_tn(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int c_dtype,
                                 int64_t ldc, int M, int N, int K, int epilogue, const float* bias,
                                 const float* row_rstd, const float* row_mean, const float* colsum,
                                 const int32_t* m_dev, void* stream) {
    int rc = check_common(A, lda, B, ldb, C, c_dtype, ldc, M, N, K, epilogue, bias, row_rstd, row_mean, colsum);
    if (rc != TASU_OK || M == 0) return rc;
    const int csz = c_dtype == TASU_F32 ? 4 : 2;
    TASU_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0), "base pointers must be 16-byte aligned");
    TASU_CHECK_ARG((lda * 2) % 16 == 0 && (ldb * 2) % 16 == 0 && (ldc * csz) % 16 == 0, "row pitches must be multiples of 16 bytes");
    // EXPERIMENTAL 16-epilogue-warp kernel for the store-heavy shallow-K shapes with bf16 output (TASU_OPT_GEMM_WIDE_EPI)
    if (option(TASU_OPT_GEMM_WIDE_EPI) != 0 && K <= 1024 && c_dtype == TASU_BF16) {
        Params pw{M, N, K, m_dev, epilogue, bias, row_rstd, row_mean, colsum};
        return launch_wide_epi(A, lda, B, ldb, C, ldc, M, N, K, pw, (cudaStream_t)stream);
    }
    // EXPERIMENTAL CTA-pair mode (off unless TASU_OPT_GEMM_PAIR is set; bit 0: deep-K shapes, bit 1: K <= 1024),
    // for problems with more than one 128-row tile
    const int pair_opt = option(TASU_OPT_GEMM_PAIR);
    const bool pair = (pair_opt & (K > 1024 ? 1 : 2)) != 0 && M > BM && sm_count() >= 2;
    CUtensorMap ma, mb, mc;
    rc = make_map(&ma, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, BM, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
    rc = make_map(&mb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldb, pair ? BN / 2 : BN, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
    rc = make_map(&mc, C, c_dtype == TASU_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, csz,
                  M, N, ldc, BM, c_dtype == TASU_F32 ? 32 : 64, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    Params p{M, N, K, m_dev, epilogue, bias, row_rstd, row_mean, colsum};
    if (pair) {
        const int pair_tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
        int clusters = sm_count() / 2;
        if (clusters > pair_tiles) clusters = pair_tiles;
        cudaStream_t pst = (cudaStream_t)stream;
        rc = c_dtype == TASU_BF16 ? launch_pair_epi<true>(epilogue, K <= 1024, clusters, pst, ma, mb, mc, p)
                                  : launch_pair_epi<false>(epilogue, K <= 1024, clusters, pst, ma, mb, mc, p);
        if (rc) return rc;
        TASU_CHECK_LAUNCH();
        return TASU_OK;
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int grid = sm_count();
    if (grid > tiles) grid = tiles;
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_dispatch(c_dtype == TASU_BF16, epilogue, K <= 1024, grid, st, ma, mb, mc, p);
    if (rc) return rc;
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_gemm_bf16_tn_simt(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int c_dtype,
                                      int64_t ldc, int M, int N, int K, int epilogue, const float* bias,
                                      const float* row_rstd, const float* row_mean, const float* colsum, void* stream) {
    int rc = check_common(A, lda, B, ldb, C, c_dtype, ldc, M, N, K, epilogue, bias, row_rstd, row_mean, colsum);
    if (rc != TASU_OK || M == 0) return rc;
    Params p{M, N, K, nullptr, epilogue, bias, row_rstd, row_mean, colsum};
    dim3 grid((N + 15) / 16, (M + 15) / 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (c_dtype == TASU_F32)
        gemm_simt_kernel<float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)B, ldb, (float*)C, ldc, p);
    else
        gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)B, ldb, (__nv_bfloat16*)C, ldc, p);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

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
    // up to 32 partials per frame (8 vocabulary splits x 4 column quarters of the wide kernel; the default kernel uses
    // 16 = 8 splits x 2 halves) x (max, sum, sum2, arg) + blank logits
    return rows * 32 * 16 + rows * 4 + 256;
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
    // EXPERIMENTAL CTA-pair mode (TASU_OPT_GEMM_PAIR bit 2): 256 frames per work item, clusters of two CTAs
    const bool pair = (option(TASU_OPT_GEMM_PAIR) & 4) != 0 && K <= 8 * BK && M > BM && sm_count() >= 2;
    if (pair) {
        rc = make_map(&mb, w_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, V, K, ldw, BN / 2, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        if (rc) return rc;
    }
    const int m_tiles = pair ? (M + 2 * BM - 1) / (2 * BM) : (M + BM - 1) / BM, n_tiles = (V + BN - 1) / BN;
    int grid = pair ? sm_count() / 2 : sm_count();           // pair mode: clusters
    StatsParams p{};
    p.M = M; p.N = V; p.K = K; p.blank = blank_id; p.bias = bias;
    pick_splits(m_tiles, n_tiles, grid, &p.splits, &p.nt_per);
    if (grid > m_tiles * p.splits) grid = m_tiles * p.splits;
    // EXPERIMENTAL 16-epilogue-warp kernel (TASU_OPT_STATS_WIDE): four partials per split instead of two
    const bool wide = option(TASU_OPT_STATS_WIDE) != 0 && !pair && K <= 8 * BK;
    const int64_t cap = wide ? 32 : 16;                      // partial slots per frame in the workspace layout
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
    const bool prefetch = (option(TASU_OPT_EPI_PREFETCH) & 2) != 0;
    if (wide) {
        static std::once_flag once_wide;
        static cudaError_t attr_err_wide = cudaSuccess;
        std::call_once(once_wide, [&] {
            attr_err_wide = cudaFuncSetAttribute(ctc_stats_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatsSmemBytesARes);
        });
        TASU_CHECK_CUDA(attr_err_wide);
        ctc_stats_wide_kernel<<<grid, kWideThreads, kStatsSmemBytesARes, st>>>(ma, mb, p);
        TASU_CHECK_LAUNCH();
        const int64_t frames_w = (int64_t)B * T;
        ctc_stats_combine_wide_kernel<<<(unsigned)((frames_w + 255) / 256), 256, 0, st>>>(p, B, T, n_prefix, argmax, x_blank, row_max,
                                                                                      row_sumexp, row_sumexp2);
        TASU_CHECK_LAUNCH();
        return TASU_OK;
    }
    if (prefetch && !pair && K <= 8 * BK) {
        auto kern = ctc_stats_kernel<true, false, true>;
        static std::once_flag once_pf;
        static cudaError_t attr_err_pf = cudaSuccess;
        std::call_once(once_pf, [&] {
            attr_err_pf = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatsSmemBytesARes);
        });
        TASU_CHECK_CUDA(attr_err_pf);
        kern<<<grid, kStatsThreads, kStatsSmemBytesARes, st>>>(ma, mb, p);
    } else if (pair) {
        static_assert(kStatsSmemBytesPair <= 227 * 1024, "pair-mode shared memory exceeds the 227 KB a CTA can opt into");
        auto kern = ctc_stats_kernel<true, true>;
        static std::once_flag once_pair;
        static cudaError_t attr_err_pair = cudaSuccess;
        std::call_once(once_pair, [&] {
            attr_err_pair = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatsSmemBytesPair);
        });
        TASU_CHECK_CUDA(attr_err_pair);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2u * (unsigned)grid);
        cfg.blockDim = dim3(kStatsThreads);
        cfg.dynamicSmemBytes = kStatsSmemBytesPair;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        TASU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, p));
    } else if (K <= 8 * BK) ctc_stats_kernel<true><<<grid, kStatsThreads, kStatsSmemBytesARes, st>>>(ma, mb, p);
    else ctc_stats_kernel<false><<<grid, kStatsThreads, kStatsSmemBytes, st>>>(ma, mb, p);
    TASU_CHECK_LAUNCH();
    const int64_t frames = (int64_t)B * T;
    ctc_stats_combine_kernel<<<(unsigned)((frames + 255) / 256), 256, 0, st>>>(p, B, T, n_prefix, argmax, x_blank, row_max, row_sumexp, row_sumexp2);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
