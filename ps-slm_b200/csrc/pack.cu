// Cross-rank packing of the compressed sequences (BASELINE.json configs[3]): after the all-gather of the
// per-rank compressed lengths and of the packed projector rows, every rank picks the utterances it will splice
// (dist.length_grouped_partition) straight out of the flat all-gather buffer.  Utterance i lives on rank i % W at local
// index i / W — the reference's sample sharding (Multitask/dataset/speech_dataset_large.py:80-91) — and rank r's rows
// occupy the slab [r * slab_rows, r * slab_rows + sum of its lengths) of the buffer.
//
// tasu_packed_select: (1) one CTA turns the gathered lengths into per-utterance source rows (exclusive scan of every
// rank's lengths) and destination offsets (exclusive scan over the selection); (2) one CTA per selected utterance
// copies its rows with 128-bit loads.  No index vector is built on the host and none is shipped to the device.
#include "common.cuh"

namespace tasu {

__global__ void __launch_bounds__(1024)
packed_offsets_kernel(const int64_t* __restrict__ all_lens, int W, int b_max, int64_t n_global, int64_t slab_rows,
                      const int32_t* __restrict__ sel, int n_sel, int32_t* __restrict__ local_off /*[W * b_max]*/,
                      int32_t* __restrict__ src_row, int32_t* __restrict__ dst_row, int64_t* __restrict__ out_lens,
                      int32_t* __restrict__ total) {
    __shared__ int scratch[33];
    // exclusive scan of every rank's lengths (utterances beyond n_global do not exist: length 0)
    for (int r = 0; r < W; ++r) {
        int carry = 0;
        for (int j0 = 0; j0 < b_max; j0 += blockDim.x) {
            const int j = j0 + threadIdx.x;
            const bool live = j < b_max && (int64_t)j * W + r < n_global;
            const int v = live ? (int)all_lens[(int64_t)r * b_max + j] : 0;
            int tot;
            const int ex = block_excl_scan_i(v, scratch, &tot);
            if (j < b_max) local_off[r * b_max + j] = carry + ex;
            carry += tot;
        }
    }
    __syncthreads();
    int carry = 0;
    for (int k0 = 0; k0 < n_sel; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        int v = 0, r = 0, j = 0;
        if (k < n_sel) {
            const int i = sel != nullptr ? sel[k] : k;
            r = i % W; j = i / W;
            v = (i >= 0 && (int64_t)i < n_global && j < b_max) ? (int)all_lens[(int64_t)r * b_max + j] : 0;
        }
        int tot;
        const int ex = block_excl_scan_i(v, scratch, &tot);
        if (k < n_sel) {
            src_row[k] = (int32_t)(r * slab_rows + local_off[r * b_max + j]);
            dst_row[k] = carry + ex;
            out_lens[k] = v;
        }
        carry += tot;
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256)
packed_copy_kernel(const uint8_t* __restrict__ src, int64_t src_stride_bytes, const int32_t* __restrict__ src_row,
                   const int32_t* __restrict__ dst_row, const int64_t* __restrict__ out_lens, int n_sel, int row_bytes,
                   int64_t max_rows, uint8_t* __restrict__ dst, int64_t dst_stride_bytes) {
    for (int k = blockIdx.x; k < n_sel; k += gridDim.x) {
        const int n = (int)out_lens[k];
        const int64_t s0 = src_row[k], d0 = dst_row[k];
        const int vec = row_bytes / 16;
        for (int f = 0; f < n; ++f) {
            if (d0 + f >= max_rows) break;                         // the destination is sized by the caller
            const uint4* s = reinterpret_cast<const uint4*>(src + (s0 + f) * src_stride_bytes);
            uint4* d = reinterpret_cast<uint4*>(dst + (d0 + f) * dst_stride_bytes);
            for (int c = threadIdx.x; c < vec; c += blockDim.x) st_stream_u4(d + c, ld_stream_u4(s + c));
        }
    }
}

}  // namespace tasu

using namespace tasu;

extern "C" int tasu_packed_select(const void* flat, int dtype, int64_t flat_row_stride, int64_t slab_rows, int H,
                                  const int64_t* all_lens, int W, int b_max, int64_t n_global, const int32_t* sel,
                                  int n_sel, void* out, int64_t out_row_stride, int64_t max_rows, int64_t* out_lens,
                                  int32_t* total, int32_t* workspace /*[W*b_max + 2*n_sel] int32*/, void* stream) {
    TASU_CHECK_ARG(W > 0 && b_max >= 0 && n_global >= 0 && n_sel >= 0 && H > 0 && slab_rows >= 0 && max_rows >= 0, "shape");
    TASU_CHECK_ARG(dtype == TASU_F32 || dtype == TASU_BF16, "dtype");
    TASU_CHECK_ARG(n_global <= (int64_t)W * b_max, "n_global exceeds W * b_max");
    TASU_CHECK_ARG((int64_t)W * slab_rows < (1LL << 31), "all-gather buffer too large for int32 row indices");
    TASU_CHECK_ARG(total != nullptr, "null total");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_sel == 0) { TASU_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(int32_t), st)); return TASU_OK; }
    TASU_CHECK_ARG(flat && all_lens && out && out_lens && workspace, "null pointer");
    const int esz = dtype == TASU_F32 ? 4 : 2;
    TASU_CHECK_ARG(((int64_t)H * esz) % 16 == 0 && (flat_row_stride * esz) % 16 == 0 && (out_row_stride * esz) % 16 == 0 &&
                   (uintptr_t)flat % 16 == 0 && (uintptr_t)out % 16 == 0, "rows must be multiples of 16 bytes and 16-byte aligned");
    int32_t* local_off = workspace;
    int32_t* src_row = workspace + (int64_t)W * b_max;
    int32_t* dst_row = src_row + n_sel;
    packed_offsets_kernel<<<1, 1024, 0, st>>>(all_lens, W, b_max, n_global, slab_rows, sel, n_sel, local_off, src_row, dst_row,
                                             out_lens, total);
    TASU_CHECK_LAUNCH();
    const int grid = n_sel < sm_count() * 8 ? n_sel : sm_count() * 8;
    packed_copy_kernel<<<grid, 256, 0, st>>>((const uint8_t*)flat, flat_row_stride * esz, src_row, dst_row, out_lens, n_sel,
                                            H * esz, max_rows, (uint8_t*)out, out_row_stride * esz);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
