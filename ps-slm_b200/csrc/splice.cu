// Step 4: splice compressed audio embeddings into the prompt-token embeddings —
// Multitask/model/ps-slm.py:765-871 (_merge_input_ids_with_audio_features), with the
// embed_tokens gather of ps-slm.py:525,654 optionally fused.
//
// The reference plans with ~25 ATen calls and ≥5 host syncs and then makes three full passes
// over [B, S', H].  Here: three tiny integer kernels produce per-TOKEN tables (new position,
// number of text tokens before it, number of audio slots before it) from which every
// destination row finds its source with one binary search, and ONE vectorised gather/scatter
// pass writes embeddings, mask, labels, position ids and final ids.
//
// Token j of row b owns the destination range (new_pos[j-1], new_pos[j]] (length = its
// placeholder count, :805-808).  Text tokens (not <speech>, mask == 1, :797-799) own exactly one
// position and are copied there (:833-840).  Every other position is an audio slot iff it lies
// in the non-pad span (:842-859); slots are numbered row-major over the batch (:867-869).
#include "common.cuh"

namespace tasu {

enum { RS_NSPEECH = 0, RS_NPAD = 1, RS_FIRST0 = 2, RS_LAST0 = 3, RS_NTEXT = 4, RS_TOT = 5, RS_SLOTS = 6, RS_WORDS = 8 };

__device__ __forceinline__ int mask_at(const void* m, int dtype, int64_t i) {
    // returns 1 for ==1, 0 for ==0, 2 for anything else (neither text nor pad in the reference)
    if (dtype == 0) { const uint8_t v = reinterpret_cast<const uint8_t*>(m)[i]; return v == 0 ? 0 : 1; }
    const int64_t v = reinterpret_cast<const int64_t*>(m)[i];
    return v == 0 ? 0 : (v == 1 ? 1 : 2);
}

__global__ void __launch_bounds__(256)
splice_rowstat_kernel(const int64_t* __restrict__ ids, const void* __restrict__ mask, int mdt, int S,
                      int64_t speech, int32_t* __restrict__ rowstat) {
    __shared__ int scratch[33];
    const int b = blockIdx.x;
    int nsp = 0, npad = 0, ntext = 0;
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const int64_t i = (int64_t)b * S + j;
        const bool sp = ids[i] == speech;
        const int m = mask_at(mask, mdt, i);
        nsp += sp;
        npad += (m == 0);
        ntext += (!sp && m == 1);
    }
    nsp = block_sum_i(nsp, scratch);
    npad = block_sum_i(npad, scratch);
    ntext = block_sum_i(ntext, scratch);
    if (threadIdx.x == 0) {
        int32_t* rs = rowstat + (int64_t)b * RS_WORDS;
        rs[RS_NSPEECH] = nsp;
        rs[RS_NPAD] = npad;
        rs[RS_FIRST0] = S > 0 ? (mask_at(mask, mdt, (int64_t)b * S) == 0) : 0;
        rs[RS_LAST0] = S > 0 ? (mask_at(mask, mdt, (int64_t)b * S + S - 1) == 0) : 0;
        rs[RS_NTEXT] = ntext;
        rs[RS_TOT] = 0; rs[RS_SLOTS] = 0; rs[7] = 0;
    }
}

// padding side of the whole batch (ps-slm.py:771-785); both sides → treated as left, error flagged by header
__device__ __forceinline__ bool batch_left_padding(const int32_t* __restrict__ rowstat, int B, int* scratch, bool* both) {
    int l = 0, r = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        l |= rowstat[(int64_t)i * RS_WORDS + RS_FIRST0];
        r |= rowstat[(int64_t)i * RS_WORDS + RS_LAST0];
    }
    l = block_max_i(l, scratch);
    r = block_max_i(r, scratch);
    bool left = true;
    *both = false;
    if (B > 1) {
        if (l && r) *both = true;
        else if (!l && r) left = false;
    }
    return left;
}

// per-batch totals: slot base of every row, audio offsets, S', padding side, error words; one CTA (any block size)
__device__ __forceinline__ void splice_header_body(const int32_t* __restrict__ rowstat, const int64_t* __restrict__ num_audio,
                                                   int n_audio, int64_t div_k, int B, int64_t* __restrict__ header,
                                                   int32_t* __restrict__ slot_base, int32_t* __restrict__ audio_off,
                                                   int* scratch) {
    bool both;
    const bool left = batch_left_padding(rowstat, B, scratch, &both);
    int carry = 0, mx = 0, nsp = 0;
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        const int v = b < B ? __ldcg(rowstat + (int64_t)b * RS_WORDS + RS_SLOTS) : 0;   // written by other CTAs when fused
        int total;
        const int excl = block_excl_scan_i(v, scratch, &total);
        if (b < B) {
            slot_base[b] = carry + excl;
            mx = max(mx, __ldcg(rowstat + (int64_t)b * RS_WORDS + RS_TOT));
            nsp += rowstat[(int64_t)b * RS_WORDS + RS_NSPEECH];
        }
        carry += total;
    }
    const int total_slots = carry;
    mx = block_max_i(mx, scratch);
    nsp = block_sum_i(nsp, scratch);
    int acarry = 0;
    for (int a0 = 0; a0 < n_audio; a0 += blockDim.x) {
        const int a = a0 + threadIdx.x;
        int v = 0;
        if (a < n_audio) { int64_t m = num_audio[a] / div_k; v = (int)(m < 0 ? 0 : m); }
        int total;
        const int excl = block_excl_scan_i(v, scratch, &total);
        if (a < n_audio) audio_off[a] = acarry + excl;
        acarry += total;
    }
    if (threadIdx.x == 0) {
        audio_off[n_audio] = acarry;
        header[TASU_SH_SPLICED_LEN] = mx;
        header[TASU_SH_LEFT_PADDING] = left ? 1 : 0;
        header[TASU_SH_ERR_BOTH_SIDES] = both ? 1 : 0;
        header[TASU_SH_TOTAL_SLOTS] = total_slots;
        header[TASU_SH_TOTAL_AUDIO] = acarry;
        header[TASU_SH_N_SPEECH] = nsp;
        header[6] = 0; header[7] = 0;
    }
}

struct SpliceHeaderArgs { int32_t* ticket; int64_t* header; int32_t* slot_base; int32_t* audio_off; };

__global__ void __launch_bounds__(256)
splice_plan_kernel(const int64_t* __restrict__ ids, const void* __restrict__ mask, int mdt, int B, int S,
                   int64_t speech, const int64_t* __restrict__ num_audio, int n_audio, int64_t div_k,
                   int32_t* __restrict__ rowstat, int32_t* __restrict__ new_pos, int32_t* __restrict__ text_prefix,
                   int32_t* __restrict__ slot_ord, const SpliceHeaderArgs hd) {
    __shared__ int scratch[33];
    __shared__ int s_last;
    const int b = blockIdx.x;
    bool both;
    const bool left = batch_left_padding(rowstat, B, scratch, &both);
    // global ordinal of this row's first <speech> token (row-major mask assignment, :806)
    int base = 0;
    for (int i = threadIdx.x; i < b; i += blockDim.x) base += rowstat[(int64_t)i * RS_WORDS + RS_NSPEECH];
    base = block_sum_i(base, scratch);
    const int n_pad = rowstat[(int64_t)b * RS_WORDS + RS_NPAD];

    // pass 1: placeholders → inclusive cumsum - 1, text prefix
    int c_sp = 0, c_ph = 0, c_tx = 0;
    for (int j0 = 0; j0 < S; j0 += blockDim.x) {
        const int j = j0 + threadIdx.x;
        int sp = 0, tx = 0;
        if (j < S) {
            const int64_t i = (int64_t)b * S + j;
            sp = ids[i] == speech;
            tx = (!sp && mask_at(mask, mdt, i) == 1);
        }
        int tot_sp, tot_tx, tot_ph;
        const int e_sp = block_excl_scan_i(sp, scratch, &tot_sp);
        const int e_tx = block_excl_scan_i(tx, scratch, &tot_tx);
        int ph = 0;
        if (j < S) {
            ph = 1;
            if (sp) {
                const int ord = base + c_sp + e_sp;
                int64_t m = 0;
                if (n_audio == 1) m = num_audio[0];
                else if (ord < n_audio) m = num_audio[ord];
                m = m / div_k;
                ph = (int)(m < 0 ? 0 : m);
            }
        }
        const int e_ph = block_excl_scan_i(ph, scratch, &tot_ph);
        if (j < S) {
            const int64_t i = (int64_t)b * S + j;
            new_pos[i] = c_ph + e_ph + ph - 1;
            text_prefix[i] = c_tx + e_tx;
        }
        c_sp += tot_sp; c_tx += tot_tx; c_ph += tot_ph;
    }
    const int tot = c_ph;
    __syncthreads();
    // pass 2: audio slots owned by every non-text token, restricted to the non-pad span
    const int span_lo = left ? n_pad : 0;
    const int span_hi = left ? tot : tot - n_pad;
    int c_sl = 0;
    for (int j0 = 0; j0 < S; j0 += blockDim.x) {
        const int j = j0 + threadIdx.x;
        int ns = 0;
        if (j < S) {
            const int64_t i = (int64_t)b * S + j;
            const bool sp = ids[i] == speech;
            const bool tx = (!sp && mask_at(mask, mdt, i) == 1);
            if (!tx) {
                const int q1 = new_pos[i] + 1;                                  // exclusive end
                const int q0 = (j == 0) ? 0 : new_pos[i - 1] + 1;               // inclusive start
                const int lo = max(q0, span_lo), hi = min(q1, span_hi);
                ns = max(0, hi - lo);
            }
        }
        int tot_sl;
        const int e_sl = block_excl_scan_i(ns, scratch, &tot_sl);
        if (j < S) slot_ord[(int64_t)b * S + j] = c_sl + e_sl;
        c_sl += tot_sl;
    }
    if (threadIdx.x == 0) {
        rowstat[(int64_t)b * RS_WORDS + RS_TOT] = tot;
        rowstat[(int64_t)b * RS_WORDS + RS_SLOTS] = c_sl;
    }
    if (hd.ticket == nullptr) return;
    // fused header (tasu_splice_plan_header): the last CTA to finish sees every row's totals; the ticket returns to zero
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int t = atomicAdd(hd.ticket, 1);
        s_last = (t == (int)gridDim.x - 1);
        if (s_last) { *hd.ticket = 0; __threadfence(); }
    }
    __syncthreads();
    if (!s_last) return;
    splice_header_body(rowstat, num_audio, n_audio, div_k, B, hd.header, hd.slot_base, hd.audio_off, scratch);
}

__global__ void __launch_bounds__(1024)
splice_header_kernel(const int32_t* __restrict__ rowstat, const int64_t* __restrict__ num_audio, int n_audio,
                     int64_t div_k, int B, int64_t* __restrict__ header, int32_t* __restrict__ slot_base,
                     int32_t* __restrict__ audio_off) {
    __shared__ int scratch[33];
    splice_header_body(rowstat, num_audio, n_audio, div_k, B, header, slot_base, audio_off, scratch);
}

struct Dest {           // what lands on destination position (b, p)
    int kind;           // 0 pad, 1 text, 2 audio
    int j;              // source token (text)
    int a;              // global audio ordinal (audio)
    int pos;            // position id
};

__device__ __forceinline__ Dest resolve_dest(int b, int p, int S, int Sp, bool left, const int32_t* __restrict__ rowstat,
                                             const int64_t* __restrict__ ids, const void* __restrict__ mask, int mdt,
                                             int64_t speech, const int32_t* __restrict__ new_pos,
                                             const int32_t* __restrict__ text_prefix, const int32_t* __restrict__ slot_ord,
                                             const int32_t* __restrict__ slot_base) {
    Dest d{0, 0, 0, 1};
    const int32_t* rs = rowstat + (int64_t)b * RS_WORDS;
    const int tot = rs[RS_TOT], n_pad = rs[RS_NPAD];
    const int q = p - (left ? (Sp - tot) : 0);
    if (q < 0 || q >= tot) return d;
    const int32_t* np = new_pos + (int64_t)b * S;
    int lo = 0, hi = S - 1;                      // smallest j with np[j] >= q (exists because q < tot)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (np[mid] >= q) hi = mid; else lo = mid + 1;
    }
    const int j = lo;
    const int64_t i = (int64_t)b * S + j;
    const bool sp = ids[i] == speech;
    const bool tx = (!sp && mask_at(mask, mdt, i) == 1);
    if (tx) {
        d.kind = 1; d.j = j; d.pos = text_prefix[i] + slot_ord[i];
        return d;
    }
    const int span_lo = left ? n_pad : 0, span_hi = left ? tot : tot - n_pad;
    if (q < span_lo || q >= span_hi) return d;
    const int q0 = (j == 0) ? 0 : np[j - 1] + 1;
    const int k_in = q - max(q0, span_lo);
    d.kind = 2;
    d.a = slot_base[b] + slot_ord[i] + k_in;
    d.pos = text_prefix[i] + slot_ord[i] + k_in;
    return d;
}

__device__ __forceinline__ int64_t audio_row_of(int a, int layout, int64_t max_len, int n_audio,
                                                const int32_t* __restrict__ audio_off) {
    if (a >= audio_off[n_audio]) return -1;
    if (layout == 0) return a;
    int lo = 0, hi = n_audio;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (audio_off[mid] <= a) lo = mid; else hi = mid;
    }
    return (int64_t)lo * max_len + (a - audio_off[lo]);
}

struct ScatterArgs {
    const int64_t* ids; const void* mask; int mdt; const int64_t* labels;
    int B, S, Sp, H;
    const void* text_src; int text_mode; int64_t text_stride;
    const void* audio; int audio_layout; int64_t audio_stride; int64_t audio_max_len; int n_audio;
    const int32_t* audio_perm;       // optional: audio row r is stored at row audio_perm[r] of `audio` (< 0: zero row)
    const int32_t* rowstat; const int32_t* new_pos; const int32_t* text_prefix; const int32_t* slot_ord;
    const int32_t* slot_base; const int32_t* audio_off; int left_padding;
    int64_t speech, pad_id, ignore_id;
    void* out_emb; void* out_mask; int64_t* out_labels; int64_t* out_pos; int64_t* out_ids;
};

// Row map: one THREAD per destination row resolves where the row comes from (text token / audio row / nothing) and
// writes the integer outputs and the map the backward gathers through:
//   row_src[row] = -1 (zero row) | text source row | kAudioFlag + audio row
constexpr int64_t kAudioFlag = 1LL << 62;
// one destination row: resolve, write the integer outputs and row_src[row]; returns row_src[row]
__device__ __forceinline__ int64_t rowmap_one(const ScatterArgs& a, int64_t row, int64_t* __restrict__ row_src,
                                              int32_t* __restrict__ audio_dest) {
    const int b = (int)(row / a.Sp), p = (int)(row % a.Sp);
    const bool left = a.left_padding != 0;
    const Dest d = resolve_dest(b, p, a.S, a.Sp, left, a.rowstat, a.ids, a.mask, a.mdt, a.speech, a.new_pos,
                                a.text_prefix, a.slot_ord, a.slot_base);
    int64_t src = -1, lab = a.ignore_id, fid = a.pad_id;
    int mval = 0;
    if (d.kind == 1) {
        const int64_t i = (int64_t)b * a.S + d.j;
        src = a.text_mode == 1 ? a.ids[i] : i;
        if (a.labels) lab = a.labels[i];
        fid = a.ids[i];
        mval = 1;
    } else if (d.kind == 2) {
        const int64_t ar = audio_row_of(d.a, a.audio_layout, a.audio_max_len, a.n_audio, a.audio_off);
        if (ar >= 0) {
            src = kAudioFlag + ar;
            if (audio_dest) audio_dest[ar] = (int32_t)row;
        }
        mval = 1;
    }
    row_src[row] = src;
    if (a.mdt == 0) reinterpret_cast<uint8_t*>(a.out_mask)[row] = (uint8_t)mval;
    else reinterpret_cast<int64_t*>(a.out_mask)[row] = mval;
    if (a.out_labels) a.out_labels[row] = lab;
    a.out_pos[row] = mval ? d.pos : 1;
    if (a.out_ids) a.out_ids[row] = fid;
    return src;
}

// 16-byte-chunk row copy with 4 independent loads in flight per lane (src == nullptr: zero fill)
__device__ __forceinline__ void copy_row_warp(const char* __restrict__ src, char* __restrict__ dst, int64_t nbytes, int lane) {
    const bool vec = (nbytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                     (src == nullptr || (reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
        const int nv = (int)(nbytes / 16);
        if (src) {
            for (int c = lane; c < nv; c += 128) {
                uint4 q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c + 32 * u < nv) q[u] = ld_stream_u4(reinterpret_cast<const uint4*>(src) + c + 32 * u);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c + 32 * u < nv) st_stream_u4(reinterpret_cast<uint4*>(dst) + c + 32 * u, q[u]);
            }
        } else {
            for (int c = lane; c < nv; c += 32) st_stream_u4(reinterpret_cast<uint4*>(dst) + c, make_uint4(0, 0, 0, 0));
        }
    } else if ((nbytes % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) &&
               (src == nullptr || (reinterpret_cast<uintptr_t>(src) & 3) == 0)) {     // odd pitches: 4-byte words
        const int nw = (int)(nbytes / 4);
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
        for (int c = lane; c < nw; c += 128) {
            uint32_t q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = (src && c + 32 * u < nw) ? s4[c + 32 * u] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c + 32 * u < nw) d4[c + 32 * u] = q[u];
        }
    } else {
        for (int64_t c = lane; c < nbytes; c += 32) dst[c] = src ? src[c] : 0;
    }
}

// Row map + copy in ONE launch: a warp owns kFusedRows consecutive destination rows; its first lanes resolve one row
// each (the dependent-load chains of the resolve run side by side and under the copies of the other resident warps),
// then the warp copies the rows one after the other.
constexpr int kFusedRows = 4;
template <int ESZ>
__global__ void __launch_bounds__(256)
splice_fused_kernel(ScatterArgs a, int64_t* __restrict__ row_src, int32_t* __restrict__ audio_dest) {
    const int lane = threadIdx.x & 31;
    const int64_t n_rows = (int64_t)a.B * a.Sp;
    const int64_t nbytes = (int64_t)a.H * ESZ;
    const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kFusedRows;
    if (row0 >= n_rows) return;                                             // warp-uniform
    int64_t mine = -1;
    if (lane < kFusedRows && row0 + lane < n_rows) {
        mine = rowmap_one(a, row0 + lane, row_src, audio_dest);
        if (a.audio_perm != nullptr && mine >= kAudioFlag) {               // permuted storage (grouped kept-frame layout)
            const int pr = a.audio_perm[mine - kAudioFlag];
            mine = pr >= 0 ? kAudioFlag + pr : -1;
        }
    }
#pragma unroll
    for (int r = 0; r < kFusedRows; ++r) {
        if (row0 + r >= n_rows) break;
        const int64_t sr = __shfl_sync(0xffffffffu, mine, r);
        const char* src = nullptr;
        if (sr >= kAudioFlag) src = reinterpret_cast<const char*>(a.audio) + (sr - kAudioFlag) * a.audio_stride * ESZ;
        else if (sr >= 0) src = reinterpret_cast<const char*>(a.text_src) + sr * a.text_stride * ESZ;
        copy_row_warp(src, reinterpret_cast<char*>(a.out_emb) + (row0 + r) * nbytes, nbytes, lane);
    }
}

// dst[r,:] = idx[r] >= 0 ? src[idx[r],:] : 0   (backward of the audio part of the splice)
template <int ESZ>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const void* __restrict__ src, int64_t src_stride, const int32_t* __restrict__ idx, int64_t n_rows, int H,
                   void* __restrict__ dst, int64_t dst_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += nwarp) {
        const int i = idx[r];
        const char* s = i >= 0 ? reinterpret_cast<const char*>(src) + (int64_t)i * src_stride * ESZ : nullptr;
        copy_row_warp(s, reinterpret_cast<char*>(dst) + r * dst_stride * ESZ, (int64_t)H * ESZ, lane);
    }
}

// backward of the TEXT part of the splice (ps-slm.py:833: final_embedding[b, text_to_overwrite] = inputs_embeds[b, non_audio]):
// dst[row_src[r], :] = src[r, :] for every destination row r that was copied from a text row (one-to-one, no accumulation);
// dst is zero-filled by the caller (speech / padded tokens receive no gradient).
template <int ESZ>
__global__ void __launch_bounds__(256)
scatter_text_rows_kernel(const void* __restrict__ src, int64_t src_stride, const int64_t* __restrict__ row_src, int64_t n_rows,
                         int H, void* __restrict__ dst, int64_t dst_stride, int64_t dst_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += nwarp) {
        const int64_t sr = row_src[r];
        if (sr < 0 || sr >= kAudioFlag || sr >= dst_rows) continue;
        copy_row_warp(reinterpret_cast<const char*>(src) + r * src_stride * ESZ,
                      reinterpret_cast<char*>(dst) + sr * dst_stride * ESZ, (int64_t)H * ESZ, lane);
    }
}

}  // namespace tasu

using namespace tasu;

extern "C" int tasu_splice_rowstat(const int64_t* input_ids, const void* attention_mask, int mask_dtype, int B, int S,
                                   int64_t speech_id, int32_t* rowstat, void* stream) {
    TASU_CHECK_ARG(B >= 0 && S >= 0, "B,S >= 0");
    TASU_CHECK_ARG(mask_dtype == 0 || mask_dtype == 1, "mask_dtype");
    if (B == 0) return TASU_OK;
    TASU_CHECK_ARG(input_ids && attention_mask && rowstat, "null pointer");
    splice_rowstat_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(input_ids, attention_mask, mask_dtype, S, speech_id, rowstat);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_splice_plan(const int64_t* input_ids, const void* attention_mask, int mask_dtype, int B, int S,
                                int64_t speech_id, const int64_t* num_audio, int n_audio, int64_t div_k,
                                int32_t* rowstat, int32_t* new_pos, int32_t* text_prefix, int32_t* slot_ord,
                                void* stream) {
    TASU_CHECK_ARG(B >= 0 && S >= 0 && n_audio >= 0 && div_k >= 1, "B,S,n_audio >= 0, div_k >= 1");
    TASU_CHECK_ARG(mask_dtype == 0 || mask_dtype == 1, "mask_dtype");
    if (B == 0) return TASU_OK;
    TASU_CHECK_ARG(input_ids && attention_mask && rowstat && new_pos && text_prefix && slot_ord, "null pointer");
    TASU_CHECK_ARG(n_audio == 0 || num_audio, "null num_audio");
    splice_plan_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(input_ids, attention_mask, mask_dtype, B, S, speech_id,
                                                            num_audio, n_audio, div_k, rowstat, new_pos, text_prefix,
                                                            slot_ord, SpliceHeaderArgs{});
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_splice_plan_header(const int64_t* input_ids, const void* attention_mask, int mask_dtype, int B, int S,
                                       int64_t speech_id, const int64_t* num_audio, int n_audio, int64_t div_k,
                                       int32_t* rowstat, int32_t* new_pos, int32_t* text_prefix, int32_t* slot_ord,
                                       int64_t* header, int32_t* slot_base, int32_t* audio_off, int32_t* ticket,
                                       void* stream) {
    TASU_CHECK_ARG(B >= 0 && S >= 0 && n_audio >= 0 && div_k >= 1, "B,S,n_audio >= 0, div_k >= 1");
    TASU_CHECK_ARG(mask_dtype == 0 || mask_dtype == 1, "mask_dtype");
    TASU_CHECK_ARG(header && audio_off && ticket, "null header outputs / ticket");
    if (B == 0) return tasu_splice_header(rowstat, num_audio, n_audio, div_k, 0, S, header, slot_base, audio_off, stream);
    TASU_CHECK_ARG(input_ids && attention_mask && rowstat && new_pos && text_prefix && slot_ord && slot_base, "null pointer");
    TASU_CHECK_ARG(n_audio == 0 || num_audio, "null num_audio");
    splice_plan_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(input_ids, attention_mask, mask_dtype, B, S, speech_id,
                                                            num_audio, n_audio, div_k, rowstat, new_pos, text_prefix,
                                                            slot_ord, SpliceHeaderArgs{ticket, header, slot_base, audio_off});
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_splice_header(const int32_t* rowstat, const int64_t* num_audio, int n_audio, int64_t div_k,
                                  int B, int S, int64_t* header, int32_t* slot_base, int32_t* audio_off, void* stream) {
    (void)S;
    TASU_CHECK_ARG(B >= 0 && n_audio >= 0 && div_k >= 1, "B,n_audio >= 0, div_k >= 1");
    TASU_CHECK_ARG(header && audio_off && (B == 0 || (rowstat && slot_base)), "null pointer");
    splice_header_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rowstat, num_audio, n_audio, div_k, B, header, slot_base, audio_off);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

static unsigned warp_row_grid(int64_t rows) {          // 8 warps per CTA, enough CTAs to fill the machine, grid-stride beyond
    int64_t g = (rows + 7) / 8, gmax = (int64_t)tasu::sm_count() * 16;
    if (g > gmax) g = gmax;
    return (unsigned)(g < 1 ? 1 : g);
}

extern "C" int tasu_splice_scatter(const int64_t* input_ids, const void* attention_mask, int mask_dtype,
                                   const int64_t* labels, int B, int S, int spliced_len, int H, int64_t speech_id,
                                   const void* text_src, int text_mode, int64_t text_row_stride,
                                   const void* audio_rows, int audio_layout, int64_t audio_row_stride,
                                   int64_t audio_max_len, int n_audio, int emb_dtype,
                                   const int32_t* rowstat, const int32_t* new_pos, const int32_t* text_prefix,
                                   const int32_t* slot_ord, const int32_t* slot_base, const int32_t* audio_off,
                                   int left_padding, int64_t pad_id, int64_t ignore_id,
                                   void* out_emb, void* out_mask, int64_t* out_labels, int64_t* out_pos,
                                   int64_t* out_ids, int64_t* row_src_ws, int32_t* audio_dest, void* stream) {
    return tasu_splice_scatter_perm(input_ids, attention_mask, mask_dtype, labels, B, S, spliced_len, H, speech_id, text_src,
                                    text_mode, text_row_stride, audio_rows, nullptr, audio_layout, audio_row_stride,
                                    audio_max_len, n_audio, emb_dtype, rowstat, new_pos, text_prefix, slot_ord, slot_base,
                                    audio_off, left_padding, pad_id, ignore_id, out_emb, out_mask, out_labels, out_pos,
                                    out_ids, row_src_ws, audio_dest, stream);
}

extern "C" int tasu_splice_scatter_perm(const int64_t* input_ids, const void* attention_mask, int mask_dtype,
                                        const int64_t* labels, int B, int S, int spliced_len, int H, int64_t speech_id,
                                        const void* text_src, int text_mode, int64_t text_row_stride,
                                        const void* audio_rows, const int32_t* audio_perm, int audio_layout,
                                        int64_t audio_row_stride, int64_t audio_max_len, int n_audio, int emb_dtype,
                                        const int32_t* rowstat, const int32_t* new_pos, const int32_t* text_prefix,
                                        const int32_t* slot_ord, const int32_t* slot_base, const int32_t* audio_off,
                                        int left_padding, int64_t pad_id, int64_t ignore_id,
                                        void* out_emb, void* out_mask, int64_t* out_labels, int64_t* out_pos,
                                        int64_t* out_ids, int64_t* row_src_ws, int32_t* audio_dest, void* stream) {
    TASU_CHECK_ARG(audio_perm == nullptr || audio_layout == 0, "audio_perm needs the packed audio layout");
    TASU_CHECK_ARG(B >= 0 && S > 0 && spliced_len >= 0 && H > 0, "shape");
    TASU_CHECK_ARG(mask_dtype == 0 || mask_dtype == 1, "mask_dtype");
    TASU_CHECK_ARG(text_mode == 0 || text_mode == 1, "text_mode");
    TASU_CHECK_ARG(audio_layout == 0 || audio_layout == 1, "audio_layout");
    TASU_CHECK_ARG(emb_dtype == TASU_F32 || emb_dtype == TASU_BF16, "emb_dtype");
    const int64_t rows = (int64_t)B * spliced_len;
    if (rows == 0) return TASU_OK;
    TASU_CHECK_ARG(rows < (1LL << 31), "too many rows");
    TASU_CHECK_ARG(input_ids && attention_mask && text_src && rowstat && new_pos && text_prefix && slot_ord &&
                   slot_base && audio_off && out_emb && out_mask && out_pos && row_src_ws, "null pointer");
    ScatterArgs a{};
    a.ids = input_ids; a.mask = attention_mask; a.mdt = mask_dtype; a.labels = labels;
    a.B = B; a.S = S; a.Sp = spliced_len; a.H = H;
    a.text_src = text_src; a.text_mode = text_mode; a.text_stride = text_row_stride;
    a.audio = audio_rows; a.audio_layout = audio_layout; a.audio_stride = audio_row_stride;
    a.audio_max_len = audio_max_len; a.n_audio = n_audio; a.audio_perm = audio_perm;
    a.rowstat = rowstat; a.new_pos = new_pos; a.text_prefix = text_prefix; a.slot_ord = slot_ord;
    a.slot_base = slot_base; a.audio_off = audio_off; a.left_padding = left_padding;
    a.speech = speech_id; a.pad_id = pad_id; a.ignore_id = ignore_id;
    a.out_emb = out_emb; a.out_mask = out_mask; a.out_labels = out_labels; a.out_pos = out_pos; a.out_ids = out_ids;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((rows + 8 * kFusedRows - 1) / (8 * kFusedRows));
    if (emb_dtype == TASU_F32) splice_fused_kernel<4><<<grid, 256, 0, st>>>(a, row_src_ws, audio_dest);
    else splice_fused_kernel<2><<<grid, 256, 0, st>>>(a, row_src_ws, audio_dest);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_gather_rows(const void* src, int dtype, int64_t src_row_stride, const int32_t* idx, int64_t n_rows, int H,
                                void* dst, int64_t dst_row_stride, void* stream) {
    TASU_CHECK_ARG(n_rows >= 0 && H > 0 && src_row_stride >= H && dst_row_stride >= H, "shape");
    TASU_CHECK_ARG(dtype == TASU_F32 || dtype == TASU_BF16, "dtype");
    if (n_rows == 0) return TASU_OK;
    TASU_CHECK_ARG(src && idx && dst, "null pointer");
    const unsigned grid = warp_row_grid(n_rows);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TASU_F32) gather_rows_kernel<4><<<grid, 256, 0, st>>>(src, src_row_stride, idx, n_rows, H, dst, dst_row_stride);
    else gather_rows_kernel<2><<<grid, 256, 0, st>>>(src, src_row_stride, idx, n_rows, H, dst, dst_row_stride);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_splice_text_grad(const void* grad_emb, int dtype, int64_t grad_row_stride, const int64_t* row_src,
                                     int64_t n_rows, int H, void* grad_text, int64_t text_row_stride, int64_t text_rows,
                                     void* stream) {
    TASU_CHECK_ARG(n_rows >= 0 && text_rows >= 0 && H > 0 && grad_row_stride >= H && text_row_stride >= H, "shape");
    TASU_CHECK_ARG(dtype == TASU_F32 || dtype == TASU_BF16, "dtype");
    cudaStream_t st = (cudaStream_t)stream;
    const int esz = dtype == TASU_F32 ? 4 : 2;
    if (text_rows > 0) {
        TASU_CHECK_ARG(grad_text != nullptr, "null output");
        TASU_CHECK_CUDA(cudaMemsetAsync(grad_text, 0, (size_t)text_rows * text_row_stride * esz, st));
    }
    if (n_rows == 0 || text_rows == 0) return TASU_OK;
    TASU_CHECK_ARG(grad_emb && row_src, "null pointer");
    const unsigned grid = warp_row_grid(n_rows);
    if (dtype == TASU_F32) scatter_text_rows_kernel<4><<<grid, 256, 0, st>>>(grad_emb, grad_row_stride, row_src, n_rows, H, grad_text, text_row_stride, text_rows);
    else scatter_text_rows_kernel<2><<<grid, 256, 0, st>>>(grad_emb, grad_row_stride, row_src, n_rows, H, grad_text, text_row_stride, text_rows);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}
