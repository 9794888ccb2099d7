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
//        every CTA runs the epilogue of its own 128 accumulator rows.
//        Default for deep-K problems (K > 1024, M > 128): projector GEMM-1 runs 7 % faster than with one CTA per tile
//        (profiles/r02a_ab.md); for K <= 1024 the pair kernel measured 11 % SLOWER and is not instantiated.
// Variants measured and removed in round 2 (profiles/r02a_ab.md): epilogue vectors of the next tile prefetched (no
// change for the kept-frame softmax GEMM), 16 independent epilogue warps with warp-private TMA stores (+6 % time).
// kGrouped (kept-frame softmax GEMM on the grouped layout, csrc/grouped.cu): the rows of a 2- / 4-row group are the
//        frames of one run; the epilogue adds their probabilities across adjacent lanes (warp shuffles, recursive halving)
//        and a tile stores 64 / 32 pooled rows through the tensor maps with 64- / 32-row boxes in `ga`.
template <bool kOutBf16, int kEpi, int kSt, int kGroups, int kMajor = 0, bool kPair = false, bool kGrouped = false>
__global__ void __launch_bounds__(128 + 128 * kGroups, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const Params p,
                    const __grid_constant__ typename GroupedSel<kGrouped>::type ga) {
    extern __shared__ __align__(1024) uint8_t smem[];          // 128B-swizzle atoms need 1024-byte alignment
    if ((smem_u32(smem) & 1023u) != 0) __trap();               // no static shared memory in this kernel → offset 0
    static_assert(!kPair || kMajor == 0, "pair mode: K-major operands only");
    static_assert(!kGrouped || (kEpi == TASU_EPI_SOFTMAX && kOutBf16 && !kPair && kMajor == 0), "grouped: bf16 softmax epilogue only");
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
        if constexpr (kGrouped) { prefetch_tmap(&ga.c2); prefetch_tmap(&ga.c4); }
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
        // Per-tile vectors (bias / colsum slices, this thread's row statistics) are fetched ONE TILE AHEAD into
        // registers, so the two dependent L2 round trips at the top of a tile no longer stall the epilogue warps (28 %
        // of their samples at K = 512).  Measured effect on the kernel time: none — the kept-frame softmax GEMM is bound
        // by the bytes it stores, not by its epilogue (profiles/r02g_softmax_gemm_attribution.md).
        constexpr bool kHasCol = kEpi == TASU_EPI_LNFOLD_SILU || kEpi == TASU_EPI_LNFOLD;
        constexpr bool kHasRow = kHasCol || kEpi == TASU_EPI_SOFTMAX;
        constexpr int kVecPer = BN / kEpiThreads;                  // columns of the tile this thread stages (2)
        float pf_bias[kVecPer], pf_col[kVecPer], pf_rstd = 1.f, pf_mean = 0.f;
        auto fetch_vectors = [&](int t) {
            const int fm0 = (t / n_tiles) * kTileM + (kPair ? (int)rank * BM : 0), fn0 = (t % n_tiles) * BN;
#pragma unroll
            for (int j = 0; j < kVecPer; ++j) {
                const int col = fn0 + et + j * kEpiThreads;
                pf_bias[j] = (kEpi != TASU_EPI_NONE && col < p.N) ? __ldg(p.bias + col) : 0.f;
                pf_col[j] = (kHasCol && col < p.N) ? __ldg(p.colsum + col) : 0.f;
            }
            pf_rstd = 1.f; pf_mean = 0.f;                          // consumed (negated) a tile later: no wait here
            if (kHasRow && fm0 + et < M_live) { pf_rstd = __ldg(p.row_rstd + fm0 + et); pf_mean = __ldg(p.row_mean + fm0 + et); }
        };
        {
            const int first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
            if (first < num_tiles) fetch_vectors(first);
        }
        // grouped layout: region boundaries (A rows / pooled rows), constant during the launch
        int gl_a2 = 0, gl_a4 = 0, gl_ax = 0, gl_o4 = 0, gl_ox = 0;
        if constexpr (kGrouped) {
            gl_a2 = __ldg(ga.lay + TASU_GL_A2); gl_a4 = __ldg(ga.lay + TASU_GL_A4); gl_ax = __ldg(ga.lay + TASU_GL_AX);
            gl_o4 = __ldg(ga.lay + TASU_GL_O4); gl_ox = __ldg(ga.lay + TASU_GL_OX);
        }
        TASU_TILE_LOOP {
            const int m0 = TASU_TILE_M0, n0 = (tile % n_tiles) * BN;
            // rows per group of this tile (1: one output row per accumulator row) and its first output row
            int grp_rows = 1, out0 = m0;
            float q_acc = 0.f;                                     // sum of p^2 over this thread's pooled columns
            if constexpr (kGrouped) {
                if (m0 >= gl_ax) out0 = gl_ox + (m0 - gl_ax);
                else if (m0 >= gl_a4) { grp_rows = 4; out0 = gl_o4 + ((m0 - gl_a4) >> 2); }
                else if (m0 >= gl_a2) { grp_rows = 2; out0 = gl_a2 + ((m0 - gl_a2) >> 1); }
            }
            // per-column vectors of this tile → shared memory (read back as broadcast LDS.128).
            // Safe to overwrite: every read of the previous tile happened before its last bar.sync.
            if (kEpi != TASU_EPI_NONE) {
#pragma unroll
                for (int j = 0; j < kVecPer; ++j) {                // each group stages its own copy (own barrier)
                    const int c = et + j * kEpiThreads;
                    // softmax epilogue works in the log2 domain: bias pre-scaled once per tile, not once per element
                    s_bias[c] = pf_bias[j] * (kEpi == TASU_EPI_SOFTMAX ? kLog2e : 1.f);
                    if (kHasCol) s_colsum[c] = pf_col[j];
                }
            }
            const float rstd = pf_rstd, nmean = -pf_mean;
            {
                const int next = tile + (kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x);
                if (next < num_tiles) fetch_vectors(next);         // in flight during this whole tile
            }
            // softmax: p = exp(acc + bias - max) / sum = 2^(acc*log2e + bias*log2e + rowc), rowc = -max*log2e + log2(1/sum):
            // one FADD + one FFMA + one MUFU.EX2 per element
            // grouped: 1/sum carries the 1/frames weight of the mean; a row of weight 0 (filler) yields exact zeros
            const float rowc = kEpi != TASU_EPI_SOFTMAX ? 0.f
                             : (kGrouped && !(rstd > 0.f)) ? -INFINITY
                             : fmaf(nmean, kLog2e, __log2f(fmaxf(rstd, 1e-37f)));
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
                if (kGrouped && grp_rows > 1) {
                    // The frames of a run are adjacent accumulator rows = adjacent lanes: recursive halving — after the
                    // exchange with lane^1 a lane holds the 2-frame sums of 16 columns (even lanes the lower half of the
                    // slab), after the exchange with lane^2 the 4-frame sums of 8 columns.  Every lane of a group adds
                    // the same values in an order that differs only by commutation: the result is deterministic.
                    const bool odd = (lane & 1) != 0, hi2 = (lane & 2) != 0;
                    float keep[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float recv = __shfl_xor_sync(0xffffffffu, odd ? f[k] : f[16 + k], 1);
                        keep[k] = (odd ? f[16 + k] : f[k]) + recv;
                    }
                    int col = n0 + sub * 32 + (odd ? 16 : 0);      // first column this lane keeps
                    if (grp_rows == 2) {
                        const int prow = et >> 1;
                        const uint32_t prow_addr = stg_u32 + (uint32_t)(sbuf * kStagingBytes + prow * 128);
                        const int piece = h * 4 + (odd ? 2 : 0), psw = prow & 7;
                        st_shared_u4(prow_addr + (uint32_t)((piece ^ psw) * 16),
                                     pack_bf16x2(keep[0], keep[1]), pack_bf16x2(keep[2], keep[3]),
                                     pack_bf16x2(keep[4], keep[5]), pack_bf16x2(keep[6], keep[7]));
                        st_shared_u4(prow_addr + (uint32_t)(((piece + 1) ^ psw) * 16),
                                     pack_bf16x2(keep[8], keep[9]), pack_bf16x2(keep[10], keep[11]),
                                     pack_bf16x2(keep[12], keep[13]), pack_bf16x2(keep[14], keep[15]));
                        if (n0 + BN <= p.N) {                      // only the last column tile has columns beyond N
#pragma unroll
                            for (int k = 0; k < 16; ++k) q_acc = fmaf(keep[k], keep[k], q_acc);
                        } else {
#pragma unroll
                            for (int k = 0; k < 16; ++k) if (col + k < p.N) q_acc = fmaf(keep[k], keep[k], q_acc);
                        }
                    } else {
                        float k4[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float recv = __shfl_xor_sync(0xffffffffu, hi2 ? keep[k] : keep[8 + k], 2);
                            k4[k] = (hi2 ? keep[8 + k] : keep[k]) + recv;
                        }
                        col += hi2 ? 8 : 0;
                        const int prow = et >> 2;
                        const uint32_t prow_addr = stg_u32 + (uint32_t)(sbuf * kStagingBytes + prow * 128);
                        const int piece = h * 4 + (odd ? 2 : 0) + (hi2 ? 1 : 0), psw = prow & 7;
                        st_shared_u4(prow_addr + (uint32_t)((piece ^ psw) * 16),
                                     pack_bf16x2(k4[0], k4[1]), pack_bf16x2(k4[2], k4[3]),
                                     pack_bf16x2(k4[4], k4[5]), pack_bf16x2(k4[6], k4[7]));
                        if (n0 + BN <= p.N) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) q_acc = fmaf(k4[k], k4[k], q_acc);
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) if (col + k < p.N) q_acc = fmaf(k4[k], k4[k], q_acc);
                        }
                    }
                } else if (kOutBf16) {
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
                        if constexpr (kGrouped) {
                            // 128 / 64 / 32 rows of the staging tile → pooled rows from out0
                            const CUtensorMap* cmap = grp_rows == 1 ? &tmap_c : (grp_rows == 2 ? &ga.c2 : &ga.c4);
                            tma_store_2d(cmap, gstaging + sbuf * kStagingBytes, n0 + ch * kColsPerChunk, out0);
                        } else if (!kPair || m0 < p.M) {
                            tma_store_2d(&tmap_c, gstaging + sbuf * kStagingBytes, n0 + ch * kColsPerChunk, m0);
                        }
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
            if constexpr (kGrouped) {
                if (grp_rows > 1) {
                    // sum of p^2 of a pooled row over the columns this epilogue group handled in this column tile
                    q_acc += __shfl_xor_sync(0xffffffffu, q_acc, 1);
                    if (grp_rows == 4) q_acc += __shfl_xor_sync(0xffffffffu, q_acc, 2);
                    const int64_t qrow = (int64_t)out0 + et / grp_rows - gl_a2;
                    if ((et & (grp_rows - 1)) == 0 && qrow < ga.ldq)
                        ga.q_part[((int64_t)(tile % n_tiles) * kGroups + grp) * ga.ldq + qrow] = q_acc;
                }
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
    gemm_bf16_tn_kernel<kOutBf16, kEpi, kSt, kGroups><<<grid, 128 + 128 * kGroups, smem, st>>>(ma, mb, mc, p, NoGroupedArgs{0});
    return TASU_OK;
}

// deep-K shapes: 4 smem stages, one epilogue group; shallow-K (store-heavy) shapes: 3 stages, two epilogue groups
template <bool kOutBf16, int kEpi>
static int launch_cfg(bool shallow_k, int grid, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb,
                      const CUtensorMap& mc, const Params& p) {
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

// CTA-pair mode (deep-K shapes; TASU_OPT_GEMM_PAIR = 0 turns it off): clusters of two CTAs, one 256x256 tile per cluster
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
    TASU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, p, NoGroupedArgs{0}));
    return TASU_OK;
}

template <bool kOutBf16>
static int launch_pair_epi(int epilogue, int clusters, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb,
                           const CUtensorMap& mc, const Params& p) {
#define TASU_PAIR(E) launch_pair_one<kOutBf16, E, kPairStages, 1>(clusters, st, ma, mb, mc, p)
    switch (epilogue) {
        case TASU_EPI_NONE: return TASU_PAIR(TASU_EPI_NONE);
        case TASU_EPI_BIAS: return TASU_PAIR(TASU_EPI_BIAS);
        case TASU_EPI_BIAS_SILU: return TASU_PAIR(TASU_EPI_BIAS_SILU);
        case TASU_EPI_BIAS_RELU: return TASU_PAIR(TASU_EPI_BIAS_RELU);
        case TASU_EPI_LNFOLD_SILU: return TASU_PAIR(TASU_EPI_LNFOLD_SILU);
        case TASU_EPI_LNFOLD: return TASU_PAIR(TASU_EPI_LNFOLD);
        default: return TASU_PAIR(TASU_EPI_SOFTMAX);
    }
#undef TASU_PAIR
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
    gemm_bf16_tn_kernel<false, TASU_EPI_NONE, kSt, kGroups, kMajor><<<grid, 128 + 128 * kGroups, smem, st>>>(ma, mb, mc, p, NoGroupedArgs{0});
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
    // CTA-pair mode for deep-K problems with more than one 128-row tile (TASU_OPT_GEMM_PAIR, default 1)
    const bool pair = option(TASU_OPT_GEMM_PAIR) != 0 && K > 1024 && M > BM && sm_count() >= 2;
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
        rc = c_dtype == TASU_BF16 ? launch_pair_epi<true>(epilogue, clusters, pst, ma, mb, mc, p)
                                  : launch_pair_epi<false>(epilogue, clusters, pst, ma, mb, mc, p);
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

// The kept-frame softmax GEMM on the grouped layout (csrc/grouped.cu): shallow-K shape (3 stages, two epilogue groups)
extern "C" int tasu_gemm_softmax_grouped_parts(int N) { return ((N + BN - 1) / BN) * 2; }

extern "C" int tasu_gemm_softmax_grouped(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int a_rows,
                                         int c_rows, int N, int K, const float* bias, const float* row_inv, const float* row_max,
                                         const int32_t* lay, float* q_part, int64_t ldq, void* stream) {
    TASU_CHECK_ARG(a_rows > 0 && c_rows > 0 && N > 0 && K > 0 && K <= 1024, "a_rows, c_rows, N > 0, 0 < K <= 1024");
    TASU_CHECK_ARG(lda >= K && ldb >= K && ldc >= N && ldq > 0, "leading dimension too small");
    TASU_CHECK_ARG(A && B && C && bias && row_inv && row_max && lay && q_part, "null pointer");
    TASU_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0), "base pointers must be 16-byte aligned");
    TASU_CHECK_ARG((lda * 2) % 16 == 0 && (ldb * 2) % 16 == 0 && (ldc * 2) % 16 == 0, "row pitches must be multiples of 16 bytes");
    CUtensorMap ma, mb, mc;
    GroupedArgs ga;
    int rc = make_map(&ma, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a_rows, K, lda, BM, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
    rc = make_map(&mb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldb, BN, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
    rc = make_map(&mc, C, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c_rows, N, ldc, BM, 64, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    rc = make_map(&ga.c2, C, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c_rows, N, ldc, BM / 2, 64, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    rc = make_map(&ga.c4, C, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c_rows, N, ldc, BM / 4, 64, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    ga.lay = lay; ga.q_part = q_part; ga.ldq = ldq;
    Params p{a_rows, N, K, lay + TASU_GL_A_ROWS, TASU_EPI_SOFTMAX, bias, row_inv, row_max, nullptr};
    const int tiles = ((a_rows + BM - 1) / BM) * ((N + BN - 1) / BN);
    int grid = sm_count();
    if (grid > tiles) grid = tiles;
    constexpr int kSt = 3, kGroups = 2;
    constexpr int smem = gemm_smem_bytes(kSt, kGroups);
    auto kern = gemm_bf16_tn_kernel<true, TASU_EPI_SOFTMAX, kSt, kGroups, 0, false, true>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    TASU_CHECK_CUDA(attr_err);
    kern<<<grid, 128 + 128 * kGroups, smem, (cudaStream_t)stream>>>(ma, mb, mc, p, ga);
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
