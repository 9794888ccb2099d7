// Shared pieces of the tcgen05 GEMM kernels (gemm_sm100.cu: one CTA per tile; gemm2cta_sm100.cu: CTA pairs):
// tile constants, PTX wrappers (mbarrier, TMA, tcgen05, TMEM), shared-memory / instruction descriptors, the
// kernel parameter block and the host-side tensor-map / argument-check helpers.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace tasu {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64;          // bf16: BK*2 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;
constexpr int kThreads = 256;                        // warps 0-3 control, warps 4-7 epilogue
constexpr int kEpiThreads = 128;
constexpr int kABytes = BM * BK * 2;                 // 16 KB
constexpr int kBBytes = BN * BK * 2;                 // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;       // 48 KB
constexpr int kStagingBytes = BM * 128;              // one 128-byte-wide column chunk of the C tile
constexpr int kAuxBytes = 2 * BN * 4;                // per-tile bias / colsum slices
constexpr int gemm_smem_bytes(int stages, int groups) {
    return stages * kStageBytes + 2 * groups * kStagingBytes + groups * kAuxBytes + 128 /*barriers*/;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (sm_100 UMMA format):
// bits [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4 = 1024 B
// between 8-row groups, [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MN-major (the operand is stored [K, MN] with MN contiguous), 128-byte swizzle: the canonical layout in 16-byte units is
// ((8,n),(8,k)):((1,LBO),(8,SBO)) — an atom is 8 K-rows x 128 B (64 MN elements); one TMA box [64 K-rows][64 MN] stacks
// 8 atoms along K (SBO = 1024 B) and consecutive boxes (the next 64 MN elements) are 8192 B apart (LBO).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)(8192u >> 4) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32 (bit4), A=B=bf16 (bits 7,10), both K-major (bit 15 / 16 = A / B MN-major),
// N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// the registers are passed as in/out operands so the compiler cannot schedule their consumers above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
__device__ __forceinline__ void st_shared_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ float silu_f(float z) { return z / (1.f + __expf(-z)); }
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {          // MUFU.EX2, 2 ulp, flushes denormal results to zero
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------ CTA-pair (cta_group::2) wrappers
// Two CTAs of one cluster (the two SMs of a TPC) run ONE tcgen05.mma of M = 256: CTA rank r holds rows [128r, 128r+128)
// of A and of the accumulator (its own TMEM) and rows [128r, 128r+128) of the 256-row B tile; the MMA is issued by the
// rank-0 CTA only and reads both shared memories at the same CTA-relative offsets.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {       // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same CTA-relative location in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // arrivals come from the peer CTA too
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t"
        "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA load into THIS CTA's shared memory whose completion bytes are signalled on a barrier that may live in the peer
// CTA of the pair (`bar_cluster_addr` from mapa_u32): both halves of a stage report to the MMA-issuing CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at the same CTA-relative offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 0x3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// instruction descriptor of the pair MMA: M = 256 (128 rows per CTA), N = 256
constexpr uint32_t kInstrDescPair = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
constexpr int kPairStages = 6;                              // deep K: 6 x 32 KB stages, one epilogue group
constexpr int kPairBBytes = (BN / 2) * BK * 2;              // each CTA loads half of the B tile: 16 KB
constexpr int kPairStageBytes = kABytes + kPairBBytes;      // 32 KB
constexpr int gemm_smem_bytes_pair(int stages, int groups) {
    return stages * kPairStageBytes + 2 * groups * kStagingBytes + groups * kAuxBytes + 256 /*barriers*/;
}

struct Params {
    int M, N, K;              // M = rows the tensor maps cover; the live row count may come from m_dev
    const int32_t* m_dev;     // optional device-side row count (data-dependent M without a host sync)
    int epilogue;
    const float* bias;
    const float* row_rstd;
    const float* row_mean;
    const float* colsum;
};

// extra kernel arguments of the grouped softmax GEMM (gemm_bf16_tn_kernel<..., kGrouped = true>; csrc/grouped.cu)
struct alignas(64) GroupedArgs {
    CUtensorMap c2, c4;       // the C matrix with 64- / 32-row store boxes (tiles of 2- / 4-frame groups)
    const int32_t* lay;       // layout words (TASU_GL_*)
    float* q_part;            // [n_tiles * groups][ldq] partial sums of p^2 of the pooled rows, row index relative to A2
    int64_t ldq;
};
struct NoGroupedArgs { int unused; };
template <bool kGrouped> struct GroupedSel { typedef NoGroupedArgs type; };
template <> struct GroupedSel<true> { typedef GroupedArgs type; };

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// 2-D row-major tensor [rows, cols] with row pitch `ld` elements; box = [box_rows, box_cols], 128B swizzle
static int make_map(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esz, int64_t rows, int64_t cols,
                    int64_t ld, int box_rows, int box_cols, CUtensorMapL2promotion promo) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return TASU_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return TASU_ERR_CUDA; }
    return TASU_OK;
}

static int check_common(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int c_dtype, int64_t ldc,
                        int M, int N, int K, int epilogue, const float* bias, const float* row_rstd,
                        const float* row_mean, const float* colsum) {
    TASU_CHECK_ARG(M >= 0 && N > 0 && K > 0, "M >= 0, N,K > 0");
    TASU_CHECK_ARG(c_dtype == TASU_F32 || c_dtype == TASU_BF16, "c_dtype");
    TASU_CHECK_ARG(epilogue >= TASU_EPI_NONE && epilogue <= TASU_EPI_SOFTMAX, "epilogue");
    TASU_CHECK_ARG(lda >= K && ldb >= K && ldc >= N, "leading dimension too small");
    if (M == 0) return TASU_OK;
    TASU_CHECK_ARG(A && B && C, "null pointer");
    TASU_CHECK_ARG(epilogue == TASU_EPI_NONE || bias, "bias required");
    TASU_CHECK_ARG((epilogue != TASU_EPI_LNFOLD_SILU && epilogue != TASU_EPI_LNFOLD) || (row_rstd && row_mean && colsum),
                   "LN-fold vectors required");
    TASU_CHECK_ARG(epilogue != TASU_EPI_SOFTMAX || (row_rstd && row_mean), "softmax row vectors required");
    return TASU_OK;
}

}  // namespace gemm
}  // namespace tasu
