// EXPERIMENTAL (not yet validated on a GPU; selected only by TASU_OPT_GEMM_WIDE_EPI): the shallow-K tcgen05 GEMM of
// gemm_sm100.cu (bf16 output) with 16 independent epilogue warps.
//
// Why (profiles/r01h_ncu_detail.md): the kept-frame softmax GEMM (M = kept frames, N = 25055, K = 512) is epilogue-
// bound — its MMA thread waits for drained accumulators, its operands are never late — although MUFU (32 %) and the
// issue slots (28 %) are mostly idle: the two groups of four epilogue warps meet at four named barriers per tile (staging
// hand-over to the TMA-store thread), load their bias / row vectors at the top of every 2 us tile, and there are only
// two warps per scheduler to hide fixed latencies.
//
// How: 16 epilogue warps (640 threads, <= 96 registers).  Warp w owns TMEM lane quadrant w % 4 and the 64-column
// quarter w / 4 of the 128 x 256 tile — exactly one 128-byte-wide bf16 staging row per accumulator row — so every warp
// stages ITS OWN 32 x 64 sub-tile in a private 4 KB swizzled buffer and issues its own TMA store (box 32 x 64): no
// barrier between epilogue warps at all, only __syncwarp.  The buffer is reused one tile later, long after its store has
// been read out (cp.async.bulk.wait_group.read 0 at the top of the tile).  Bias (2 values per lane, warp-private copy in
// shared memory), colsum and the row statistics of the NEXT tile are prefetched while the current tile is processed.
// The accumulator is read as 16-column tcgen05.ld slabs.  Producer and MMA roles are those of gemm_bf16_tn_kernel
// (3 x 48 KB stages).  Arithmetic per element is unchanged: results must be bit-identical to the default kernel.
#include "gemm_common.cuh"

namespace tasu {
namespace gemm {

constexpr int kWideStages = 3;
constexpr int kWideThreadsG = 640;                   // warps 0-3 control, warps 4-19 epilogue
constexpr int kWideEpiThreadsG = 512;
constexpr int kWideWarpStage = 32 * 128;             // private staging: 32 rows x 128 B (64 bf16 columns)
constexpr int kWideAux = 16 * 2 * 64 * 4;            // per warp: bias[64] + colsum[64] floats
constexpr int kWideSmem = kWideStages * kStageBytes + 16 * kWideWarpStage + kWideAux + 128;

template <int kEpi>
__global__ void __launch_bounds__(kWideThreadsG, 1)
gemm_wide_epi_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_c, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* staging = smem + kWideStages * kStageBytes;                       // [16 warps] x 4 KB (1024-byte aligned)
    float* s_aux = reinterpret_cast<float*>(staging + 16 * kWideWarpStage);    // [16 warps][bias 64 | colsum 64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 16 * kWideWarpStage + kWideAux);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWideStages;
    uint64_t* tmem_full = bars + 2 * kWideStages;
    uint64_t* tmem_empty = bars + 2 * kWideStages + kAccStages;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWideStages + 2 * kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M_live = p.m_dev != nullptr ? min(max(__ldg(p.m_dev), 0), p.M) : p.M;
    const int m_tiles = (M_live + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int k_blocks = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); prefetch_tmap(&tmap_c); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kWideStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kWideEpiThreadsG); }
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kStageBytes;
                    mbar_expect_tx(&full_bar[stage], kStageBytes);
                    tma_load_2d(&tmap_a, &full_bar[stage], sa, kb * BK, m0);
                    tma_load_2d(&tmap_b, &full_bar[stage], sa + kABytes, kb * BK, n0);
                    if (++stage == kWideStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kInstrDesc,
                                  (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == kWideStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== 16 independent epilogue warps =====================
        constexpr bool kLnFold = kEpi == TASU_EPI_LNFOLD_SILU || kEpi == TASU_EPI_LNFOLD;
        constexpr bool kRowVec = kLnFold || kEpi == TASU_EPI_SOFTMAX;
        const int wq = warp - 4;
        const int ew = wq & 3;                                     // TMEM lanes [32*ew, 32*ew+32)
        const int quarter = wq >> 2;                               // columns [64*quarter, 64*quarter+64) of the tile
        uint8_t* wstage = staging + wq * kWideWarpStage;
        float* w_bias = s_aux + wq * 128;
        float* w_colsum = w_bias + 64;
        const uint32_t srow = smem_u32(wstage) + (uint32_t)(lane * 128);
        const int sw = lane & 7;
        int acc = 0; uint32_t acc_phase = 0;

        // values of the NEXT tile: lane l holds columns 2l, 2l+1 of the warp's quarter and its own row's statistics
        float pf_b0 = 0.f, pf_b1 = 0.f, pf_c0 = 0.f, pf_c1 = 0.f, pf_rstd = 1.f, pf_mean = 0.f;
        auto prefetch_tile = [&](int t) {
            const int pm0 = (t / n_tiles) * BM, pn0 = (t % n_tiles) * BN;
            const int col = pn0 + quarter * 64 + 2 * lane;
            pf_b0 = (kEpi != TASU_EPI_NONE && col < p.N) ? __ldg(p.bias + col) : 0.f;
            pf_b1 = (kEpi != TASU_EPI_NONE && col + 1 < p.N) ? __ldg(p.bias + col + 1) : 0.f;
            pf_c0 = (kLnFold && col < p.N) ? __ldg(p.colsum + col) : 0.f;
            pf_c1 = (kLnFold && col + 1 < p.N) ? __ldg(p.colsum + col + 1) : 0.f;
            pf_rstd = 1.f; pf_mean = 0.f;
            const int prow = pm0 + ew * 32 + lane;
            if (kRowVec && prow < M_live) { pf_rstd = __ldg(p.row_rstd + prow); pf_mean = __ldg(p.row_mean + prow); }
        };
        if ((int)blockIdx.x < num_tiles) prefetch_tile((int)blockIdx.x);

        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
            const int c0 = n0 + quarter * 64;                      // first column of this warp's sub-tile
            // the store of the previous tile has read this warp's staging buffer (and every lane has finished with the
            // previous tile's bias copy: the __syncwarp below orders the lanes)
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            if (kEpi != TASU_EPI_NONE) {
                const float sc = kEpi == TASU_EPI_SOFTMAX ? kLog2e : 1.f;
                *reinterpret_cast<float2*>(w_bias + 2 * lane) = make_float2(pf_b0 * sc, pf_b1 * sc);
                if (kLnFold) *reinterpret_cast<float2*>(w_colsum + 2 * lane) = make_float2(pf_c0, pf_c1);
            }
            const float rstd = pf_rstd, nmean = -pf_mean;
            {
                const int nxt = tile + (int)gridDim.x;             // in flight while this tile is processed
                if (nxt < num_tiles) prefetch_tile(nxt);
            }
            __syncwarp();                                          // bias copy visible to the whole warp
            const float rowc = kEpi == TASU_EPI_SOFTMAX ? fmaf(nmean, kLog2e, __log2f(fmaxf(rstd, 1e-37f))) : 0.f;
            const bool live = c0 < p.N;                            // warp-uniform: the quarter lies inside the matrix
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + quarter * 64);

            // one 16-column slab (s = 0..3 inside the quarter): epilogue math in registers, swizzled st.shared
            auto process = [&](uint32_t (&v)[16], int s) {
                if (!live) return;
                const float4* b4 = reinterpret_cast<const float4*>(w_bias + s * 16);
                const float4* c4 = reinterpret_cast<const float4*>(w_colsum + s * 16);
                float f[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float x[4] = {__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                  __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])};
                    if (kEpi != TASU_EPI_NONE) {
                        const float4 bb = b4[q];
                        const float b[4] = {bb.x, bb.y, bb.z, bb.w};
                        if (kLnFold) {
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
#pragma unroll
                for (int q = 0; q < 2; ++q)                        // 16 columns = 32 bytes = 16-byte pieces 2s, 2s+1
                    st_shared_u4(srow + (uint32_t)((((2 * s + q) ^ sw)) * 16),
                                 pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                                 pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
            };

            uint32_t va[16], vb[16];
            tmem_ld16(t_row, va);
            tmem_ld_wait16(va);
            tmem_ld16(t_row + 16u, vb);
            process(va, 0);
            tmem_ld_wait16(vb);
            tmem_ld16(t_row + 32u, va);
            process(vb, 1);
            tmem_ld_wait16(va);
            tmem_ld16(t_row + 48u, vb);
            process(va, 2);
            tmem_ld_wait16(vb);
            tc_fence_before();                                     // every tcgen05.ld of this thread has completed
            mbar_arrive(&tmem_empty[acc]);
            process(vb, 3);

            fence_proxy_async_smem();                              // this lane's st.shared before the TMA store reads them
            __syncwarp();
            if (lane == 0) {
                if (live && m0 + ew * 32 < p.M) tma_store_2d(&tmap_c, wstage, c0, m0 + ew * 32);   // box inside the matrix
                tma_store_commit();
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
}

template <int kEpi>
static int launch_wide_one(int grid, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                           const Params& p) {
    static_assert(kWideSmem <= 227 * 1024, "shared memory exceeds the 227 KB a CTA can opt into");
    auto kern = gemm_wide_epi_kernel<kEpi>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWideSmem); });
    TASU_CHECK_CUDA(attr_err);
    kern<<<grid, kWideThreadsG, kWideSmem, st>>>(ma, mb, mc, p);
    return TASU_OK;
}

// bf16 output, K-major operands, any epilogue; C tensor map with a [32 rows x 64 columns] box.  Called by
// tasu_gemm_bf16_tn when TASU_OPT_GEMM_WIDE_EPI is set (gemm_sm100.cu).
int launch_wide_epi(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int M, int N, int K,
                    const Params& p, cudaStream_t st) {
    CUtensorMap ma, mb, mc;
    int rc = make_map(&ma, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, BM, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
    rc = make_map(&mb, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldb, BN, BK, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
    rc = make_map(&mc, C, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, ldc, 32, 64, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int grid = sm_count();
    if (grid > tiles) grid = tiles;
    switch (p.epilogue) {
        case TASU_EPI_NONE: rc = launch_wide_one<TASU_EPI_NONE>(grid, st, ma, mb, mc, p); break;
        case TASU_EPI_BIAS: rc = launch_wide_one<TASU_EPI_BIAS>(grid, st, ma, mb, mc, p); break;
        case TASU_EPI_BIAS_SILU: rc = launch_wide_one<TASU_EPI_BIAS_SILU>(grid, st, ma, mb, mc, p); break;
        case TASU_EPI_BIAS_RELU: rc = launch_wide_one<TASU_EPI_BIAS_RELU>(grid, st, ma, mb, mc, p); break;
        case TASU_EPI_LNFOLD_SILU: rc = launch_wide_one<TASU_EPI_LNFOLD_SILU>(grid, st, ma, mb, mc, p); break;
        case TASU_EPI_LNFOLD: rc = launch_wide_one<TASU_EPI_LNFOLD>(grid, st, ma, mb, mc, p); break;
        default: rc = launch_wide_one<TASU_EPI_SOFTMAX>(grid, st, ma, mb, mc, p); break;
    }
    if (rc) return rc;
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

}  // namespace gemm
}  // namespace tasu
