// Token-row projector: EncoderProjectorLinearSiLU (Multitask/model/projector.py:149-151) applied to
// TEXT-SIMULATED posteriors (Multitask/model/ps-slm.py:337-358, :360-409) without ever building them.
//
// Every row the simulators emit is  x = base·1 + (hot − base)·e_t  (t = token id; clean rows: hot=1, base=0;
// smoothed rows: hot = (1−α)+α/V, base = α/V; inserted hard blanks: t = blank, hot=1, base=0).  For such a row
//   mean = (hot + (V−1)·base)/V,   var = (hot² + (V−1)·base²)/V − mean²,   rstd = 1/sqrt(var + eps)
//   W1·LN(x) + b1 = a·γ_t·W1[:,t] + e·S + D,    a = rstd·(hot − base),  e = rstd·(base − mean),
//   S = W1·γ (row dots),  D = W1·β + b1
// so the 2·V·2048-flop GEMM row collapses to one column gather of W1 — fp32-exact, HBM-bound — and the backward
// to a column scatter:  with dz = dL/dz,
//   dW1[j,v] = γ_v·(P[j,v] + E_j) + β_v·db1_j,   P[j,v] = Σ_{r: t_r = v} a_r dz[r,j],  E_j = Σ_r e_r dz[r,j],
//   dγ_v = Σ_j W1[j,v]·(P[j,v] + E_j),   dβ_v = Σ_j W1[j,v]·db1_j,   db1_j = Σ_r dz[r,j].
// Rows are grouped by token on the host (tasu_host_group_tokens, a stable counting sort) so that every W1 column is
// touched once per step and all sums are deterministic.
#include "common.cuh"

#include <vector>

namespace tasu {

// S[j] = Σ_v W1[j,v]·γ_v ;  D[j] = Σ_v W1[j,v]·β_v + b1[j]   (fp32; one CTA per output feature j)
__global__ void __launch_bounds__(256)
rowdots_kernel(const float* __restrict__ w1, int64_t wstride, const float* __restrict__ gamma,
               const float* __restrict__ beta, const float* __restrict__ b1, int K, float* __restrict__ S,
               float* __restrict__ D) {
    __shared__ float red[8];
    const int j = blockIdx.x;
    const float* w = w1 + (int64_t)j * wstride;
    float s0 = 0.f, s1 = 0.f, d0 = 0.f, d1 = 0.f;
    const int bd = blockDim.x;
    int k = threadIdx.x;
    for (; k + 3 * bd < K; k += 4 * bd) {                      // 4 independent coalesced loads in flight
        const float a = w[k], b = w[k + bd], c = w[k + 2 * bd], d = w[k + 3 * bd];
        s0 = fmaf(a, gamma[k], s0); s1 = fmaf(b, gamma[k + bd], s1);
        s0 = fmaf(c, gamma[k + 2 * bd], s0); s1 = fmaf(d, gamma[k + 3 * bd], s1);
        d0 = fmaf(a, beta[k], d0); d1 = fmaf(b, beta[k + bd], d1);
        d0 = fmaf(c, beta[k + 2 * bd], d0); d1 = fmaf(d, beta[k + 3 * bd], d1);
    }
    for (; k < K; k += bd) { const float a = w[k]; s0 = fmaf(a, gamma[k], s0); d0 = fmaf(a, beta[k], d0); }
    const float s = block_sum_f(s0 + s1, red);
    const float d = block_sum_f(d0 + d1, red);
    if (threadIdx.x == 0) { S[j] = s; D[j] = d + (b1 ? b1[j] : 0.f); }
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }

// one CTA per distinct token: column W1[:,t]·γ_t staged once in shared memory, then one pass per row of the group
__global__ void __launch_bounds__(256)
tokrow_fwd_kernel(const float* __restrict__ w1, int64_t wstride, const float* __restrict__ gamma,
                  const float* __restrict__ S, const float* __restrict__ D, const int32_t* __restrict__ uniq,
                  const int32_t* __restrict__ seg_off, const int32_t* __restrict__ perm, const float* __restrict__ hot,
                  const float* __restrict__ base, int n_uniq, int V, int Hb, float eps, float* __restrict__ z,
                  __nv_bfloat16* __restrict__ h, float* __restrict__ row_a, float* __restrict__ row_e,
                  const float* __restrict__ colT) {
    extern __shared__ float col[];                              // [Hb]
    for (int u = blockIdx.x; u < n_uniq; u += gridDim.x) {
        if (colT != nullptr) {                                  // columns already gathered by tokrow_cols_kernel
            const float* c = colT + (int64_t)u * Hb;
            for (int j = threadIdx.x * 4; j < Hb; j += blockDim.x * 4)
                *reinterpret_cast<float4*>(col + j) = *reinterpret_cast<const float4*>(c + j);
        } else {
            const int v = uniq[u];
            const float g = gamma[v];
            for (int j = threadIdx.x; j < Hb; j += blockDim.x) col[j] = w1[(int64_t)j * wstride + v] * g;
        }
        __syncthreads();
        const int i0 = seg_off[u], i1 = seg_off[u + 1];
        for (int i = i0; i < i1; ++i) {
            const int r = perm[i];
            // closed-form LayerNorm statistics of a two-valued row (double: no cancellation)
            const double hv = (double)hot[r], bv = (double)base[r];
            const double mean = (hv + (double)(V - 1) * bv) / V;
            double var = (hv * hv + (double)(V - 1) * bv * bv) / V - mean * mean;
            var = var < 0 ? 0 : var;
            const double rstd = 1.0 / sqrt(var + (double)eps);
            const float a = (float)(rstd * (hv - bv)), e = (float)(rstd * (bv - mean));
            if (threadIdx.x == 0) { row_a[r] = a; row_e[r] = e; }
            float* zr = z ? z + (int64_t)r * Hb : nullptr;
            __nv_bfloat16* hr = h + (int64_t)r * Hb;
            for (int j = threadIdx.x * 8; j < Hb; j += blockDim.x * 8) {
                const float4 s0 = *reinterpret_cast<const float4*>(S + j), s1 = *reinterpret_cast<const float4*>(S + j + 4);
                const float4 d0 = *reinterpret_cast<const float4*>(D + j), d1 = *reinterpret_cast<const float4*>(D + j + 4);
                const float4 c0 = *reinterpret_cast<const float4*>(col + j), c1 = *reinterpret_cast<const float4*>(col + j + 4);
                float4 z0, z1;
                z0.x = fmaf(a, c0.x, fmaf(e, s0.x, d0.x)); z0.y = fmaf(a, c0.y, fmaf(e, s0.y, d0.y));
                z0.z = fmaf(a, c0.z, fmaf(e, s0.z, d0.z)); z0.w = fmaf(a, c0.w, fmaf(e, s0.w, d0.w));
                z1.x = fmaf(a, c1.x, fmaf(e, s1.x, d1.x)); z1.y = fmaf(a, c1.y, fmaf(e, s1.y, d1.y));
                z1.z = fmaf(a, c1.z, fmaf(e, s1.z, d1.z)); z1.w = fmaf(a, c1.w, fmaf(e, s1.w, d1.w));
                if (zr) { *reinterpret_cast<float4*>(zr + j) = z0; *reinterpret_cast<float4*>(zr + j + 4) = z1; }
                *reinterpret_cast<uint4*>(hr + j) = make_uint4(pack_bf16x2(silu_f(z0.x), silu_f(z0.y)), pack_bf16x2(silu_f(z0.z), silu_f(z0.w)),
                                                               pack_bf16x2(silu_f(z1.x), silu_f(z1.y)), pack_bf16x2(silu_f(z1.z), silu_f(z1.w)));
            }
        }
        __syncthreads();
    }
}

// row_slot[perm[i]] = u for seg_off[u] <= i < seg_off[u+1]  (one thread per sorted position, binary search)
__global__ void __launch_bounds__(256)
row_slot_kernel(const int32_t* __restrict__ seg_off, const int32_t* __restrict__ perm, int n_uniq, int64_t n_rows,
                int32_t* __restrict__ row_slot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    int lo = 0, hi = n_uniq;                                   // largest u with seg_off[u] <= i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_off[mid] <= (int)i) lo = mid; else hi = mid; }
    row_slot[perm[i]] = lo;
}

// Training forward, row pass: one WARP per row, no shared memory, no barriers.  z[r,:] = a·colT[slot(r),:] + e·S + D,
// h = bf16(silu(z)); the compact column colT[u] (8 KB) is streamed from L2 (rows of the same token re-read it there).
__global__ void __launch_bounds__(256)
tokrow_rows_kernel(const float* __restrict__ colT, const int32_t* __restrict__ row_slot, const float* __restrict__ S,
                   const float* __restrict__ D, const float* __restrict__ hot, const float* __restrict__ base, int64_t n_rows,
                   int V, int Hb, float eps, float* __restrict__ z, __nv_bfloat16* __restrict__ h, float* __restrict__ row_a,
                   float* __restrict__ row_e) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += nwarp) {
        const float* c = colT + (int64_t)row_slot[r] * Hb;
        const double hv = (double)hot[r], bv = (double)base[r];
        const double mean = (hv + (double)(V - 1) * bv) / V;
        double var = (hv * hv + (double)(V - 1) * bv * bv) / V - mean * mean;
        var = var < 0 ? 0 : var;
        const double rstd = 1.0 / sqrt(var + (double)eps);
        const float a = (float)(rstd * (hv - bv)), e = (float)(rstd * (bv - mean));
        if (lane == 0) { row_a[r] = a; row_e[r] = e; }
        float* zr = z ? z + r * Hb : nullptr;
        __nv_bfloat16* hr = h + r * Hb;
        for (int j = lane * 8; j < Hb; j += 256) {
            const float4 c0 = *reinterpret_cast<const float4*>(c + j), c1 = *reinterpret_cast<const float4*>(c + j + 4);
            const float4 s0 = *reinterpret_cast<const float4*>(S + j), s1 = *reinterpret_cast<const float4*>(S + j + 4);
            const float4 d0 = *reinterpret_cast<const float4*>(D + j), d1 = *reinterpret_cast<const float4*>(D + j + 4);
            float4 z0, z1;
            z0.x = fmaf(a, c0.x, fmaf(e, s0.x, d0.x)); z0.y = fmaf(a, c0.y, fmaf(e, s0.y, d0.y));
            z0.z = fmaf(a, c0.z, fmaf(e, s0.z, d0.z)); z0.w = fmaf(a, c0.w, fmaf(e, s0.w, d0.w));
            z1.x = fmaf(a, c1.x, fmaf(e, s1.x, d1.x)); z1.y = fmaf(a, c1.y, fmaf(e, s1.y, d1.y));
            z1.z = fmaf(a, c1.z, fmaf(e, s1.z, d1.z)); z1.w = fmaf(a, c1.w, fmaf(e, s1.w, d1.w));
            if (zr) { *reinterpret_cast<float4*>(zr + j) = z0; *reinterpret_cast<float4*>(zr + j + 4) = z1; }
            *reinterpret_cast<uint4*>(hr + j) = make_uint4(pack_bf16x2(silu_f(z0.x), silu_f(z0.y)), pack_bf16x2(silu_f(z0.z), silu_f(z0.w)),
                                                           pack_bf16x2(silu_f(z1.x), silu_f(z1.y)), pack_bf16x2(silu_f(z1.z), silu_f(z1.w)));
        }
    }
}

__device__ __forceinline__ float silu_grad(float zz) {
    const float s = 1.f / (1.f + __expf(-zz));
    return s * (1.f + zz * (1.f - s));
}

// one CTA per distinct token (grid-stride): dz = dh·silu'(z) for the rows of the group, P[u,:] = Σ a_r dz[r,:];
// db1 / E are accumulated per CTA in registers and written as per-CTA partials (part[cta][0|1][Hb]) that
// reduce_partials_kernel sums in CTA order — no atomics, bit-reproducible.  blockDim.x·8·CHUNKS must cover Hb.
template <int CHUNKS>
__global__ void __launch_bounds__(256)
tokrow_bwd_rows_kernel(const float* __restrict__ dh, const float* __restrict__ z, int Hb, const int32_t* __restrict__ seg_off,
                       const int32_t* __restrict__ perm, const float* __restrict__ row_a, const float* __restrict__ row_e,
                       int n_uniq, float* __restrict__ P, float* __restrict__ part) {
    float adb[CHUNKS][8], ae[CHUNKS][8];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) { adb[c][k] = 0.f; ae[c][k] = 0.f; }
    for (int u = blockIdx.x; u < n_uniq; u += gridDim.x) {
        const int i0 = seg_off[u], i1 = seg_off[u + 1];
        float ap[CHUNKS][8];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
            for (int k = 0; k < 8; ++k) ap[c][k] = 0.f;
        for (int i = i0; i < i1; ++i) {
            const int r = perm[i];
            const float a = row_a[r], e = row_e[r];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const int j = (c * blockDim.x + threadIdx.x) * 8;
                if (j < Hb) {
                    const float* dp = dh + (int64_t)r * Hb + j;
                    const float* zp = z + (int64_t)r * Hb + j;
                    const uint4 q0 = ld_stream_u4(dp), q1 = ld_stream_u4(dp + 4), y0 = ld_stream_u4(zp), y1 = ld_stream_u4(zp + 4);
                    float d[8], zz[8];
                    d[0] = __uint_as_float(q0.x); d[1] = __uint_as_float(q0.y); d[2] = __uint_as_float(q0.z); d[3] = __uint_as_float(q0.w);
                    d[4] = __uint_as_float(q1.x); d[5] = __uint_as_float(q1.y); d[6] = __uint_as_float(q1.z); d[7] = __uint_as_float(q1.w);
                    zz[0] = __uint_as_float(y0.x); zz[1] = __uint_as_float(y0.y); zz[2] = __uint_as_float(y0.z); zz[3] = __uint_as_float(y0.w);
                    zz[4] = __uint_as_float(y1.x); zz[5] = __uint_as_float(y1.y); zz[6] = __uint_as_float(y1.z); zz[7] = __uint_as_float(y1.w);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float dz = d[k] * silu_grad(zz[k]);
                        adb[c][k] += dz;
                        ae[c][k] = fmaf(e, dz, ae[c][k]);
                        ap[c][k] = fmaf(a, dz, ap[c][k]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int j = (c * blockDim.x + threadIdx.x) * 8;
            if (j < Hb) {
                float* pp = P + (int64_t)u * Hb + j;
                *reinterpret_cast<float4*>(pp) = make_float4(ap[c][0], ap[c][1], ap[c][2], ap[c][3]);
                *reinterpret_cast<float4*>(pp + 4) = make_float4(ap[c][4], ap[c][5], ap[c][6], ap[c][7]);
            }
        }
    }
    float* mine = part + (int64_t)blockIdx.x * 2 * Hb;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int j = (c * blockDim.x + threadIdx.x) * 8;
        if (j < Hb) {
            *reinterpret_cast<float4*>(mine + j) = make_float4(adb[c][0], adb[c][1], adb[c][2], adb[c][3]);
            *reinterpret_cast<float4*>(mine + j + 4) = make_float4(adb[c][4], adb[c][5], adb[c][6], adb[c][7]);
            *reinterpret_cast<float4*>(mine + Hb + j) = make_float4(ae[c][0], ae[c][1], ae[c][2], ae[c][3]);
            *reinterpret_cast<float4*>(mine + Hb + j + 4) = make_float4(ae[c][4], ae[c][5], ae[c][6], ae[c][7]);
        }
    }
}

// out0[c] = Σ_p part[p][0][c], out1[c] = Σ_p part[p][1][c] in a fixed order: CTA = 32 columns x 32 warps, warp w sums
// p = w, w+32, ... (all loads independent), then the 32 warp sums are combined in warp order.
__global__ void __launch_bounds__(1024)
reduce_partials_kernel(const float* __restrict__ part, int n_parts, int C, float* __restrict__ out0, float* __restrict__ out1,
                       const float* __restrict__ add1) {
    __shared__ float s[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float a0 = 0.f, a1 = 0.f;
    if (c < 2 * C) {
        int p = warp;
        for (; p + 32 < n_parts; p += 64) { a0 += part[(int64_t)p * 2 * C + c]; a1 += part[(int64_t)(p + 32) * 2 * C + c]; }
        if (p < n_parts) a0 += part[(int64_t)p * 2 * C + c];
    }
    s[warp][lane] = a0 + a1;
    __syncthreads();
    if (warp == 0 && c < 2 * C) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < 32; ++w) r += s[w][lane];
        if (c < C) out0[c] = r; else out1[c - C] = r + (add1 ? add1[c - C] : 0.f);
    }
}

// Training forward, dense pass over W1 (every sector used once): CTA = 32 vocabulary columns x all output features.
//   colT[slot(v), j] = W1[j,v]·γ_v  for the tokens present in the batch (compact, j contiguous),
//   part[cta][0][j]  = Σ_{v in CTA} W1[j,v]·γ_v,   part[cta][1][j] = Σ_{v in CTA} W1[j,v]·β_v   (→ S, D by reduce_partials)
__global__ void __launch_bounds__(256)
tokrow_cols_kernel(const float* __restrict__ w1, int64_t wstride, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const int32_t* __restrict__ slot, int Hb, int V,
                   float* __restrict__ colT, float* __restrict__ part) {
    __shared__ float tile[2][64][33];                          // [0] = W1·γ, [1] = W1·β for 64 features x 32 columns
    __shared__ int s_slot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v0 = blockIdx.x * 32;
    if (threadIdx.x < 32) s_slot[threadIdx.x] = (v0 + threadIdx.x < V) ? slot[v0 + threadIdx.x] : -1;
    const int v = v0 + lane;
    const bool vok = v < V;
    const float gm = vok ? gamma[v] : 0.f, bt = vok ? beta[v] : 0.f;
    __syncthreads();
    int sl[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) sl[q] = s_slot[warp * 4 + q];
    float* mine = part + (int64_t)blockIdx.x * 2 * Hb;
    const float* wp = w1 + (int64_t)(warp * 8) * wstride + v;
    for (int j0 = 0; j0 < Hb; j0 += 64) {
        float wv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) wv[q] = (vok && j0 + warp * 8 + q < Hb) ? wp[(int64_t)(j0 + q) * wstride] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) { tile[0][warp * 8 + q][lane] = wv[q] * gm; tile[1][warp * 8 + q][lane] = wv[q] * bt; }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (sl[q] >= 0) {
                float* dst = colT + (int64_t)sl[q] * Hb + j0;
                if (j0 + lane < Hb) dst[lane] = tile[0][lane][warp * 4 + q];
                if (j0 + 32 + lane < Hb) dst[32 + lane] = tile[0][32 + lane][warp * 4 + q];
            }
        }
        if (threadIdx.x < 128) {                               // row sums over the CTA's 32 columns (fixed order)
            const int which = threadIdx.x >> 6, jl = threadIdx.x & 63;
            const float* row = tile[which][jl];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < 32; k += 4) { a0 += row[k]; a1 += row[k + 1]; a2 += row[k + 2]; a3 += row[k + 3]; }
            if (j0 + jl < Hb) mine[(int64_t)which * Hb + j0 + jl] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
slot_scatter_kernel(const int32_t* __restrict__ uniq, int n_uniq, int32_t* __restrict__ slot) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_uniq) slot[uniq[u]] = u;
}

// dW1 / dγ / dβ in one pass over W1: CTA = 32 vocabulary columns x ALL output features (so dγ/dβ need no atomics),
// 64 features per iteration; the compact P rows ([n_uniq, Hb], j contiguous) are transposed through shared memory
// into the [j, v] orientation of W1 / dW1.  8 independent loads per thread in flight in each phase.
__global__ void __launch_bounds__(256, 6)
tokrow_wgrad_finish_kernel(const float* __restrict__ P, const int32_t* __restrict__ slot, const float* __restrict__ w1,
                           int64_t wstride, const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ E, const float* __restrict__ db1, int Hb, int V,
                           float* __restrict__ dw1, int64_t dstride, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    // 6 CTAs per SM (<= 40 registers): the ceil(V/32) = 783 CTAs of the real shape are resident in ONE wave
    extern __shared__ float s_vec[];                            // E[Hb] | db1[Hb]
    __shared__ float tile[32][33];
    __shared__ float s_ag[8][32], s_ab[8][32];
    __shared__ int s_slot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v0 = blockIdx.x * 32;
    if (threadIdx.x < 32) s_slot[threadIdx.x] = (v0 + threadIdx.x < V) ? slot[v0 + threadIdx.x] : -1;
    for (int j = threadIdx.x; j < Hb; j += blockDim.x) { s_vec[j] = E[j]; s_vec[Hb + j] = db1[j]; }
    const int v = v0 + lane;
    const bool vok = v < V;
    const float gm = vok ? gamma[v] : 0.f, bt = vok ? beta[v] : 0.f;
    float ag = 0.f, ab = 0.f;
    __syncthreads();
    const float* wp = w1 + (int64_t)(warp * 4) * wstride + v;
    float* dp = dw1 + (int64_t)(warp * 4) * dstride + v;
    for (int j0 = 0; j0 < Hb; j0 += 32) {
        float pv[4], wv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                          // P rows of this warp's 4 columns → tile[v][j]
            const int sl = s_slot[warp * 4 + q];
            pv[q] = (sl >= 0 && j0 + lane < Hb) ? P[(int64_t)sl * Hb + j0 + lane] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)                            // W1 loads do not depend on the tile: issue them now
            wv[q] = (vok && j0 + warp * 4 + q < Hb) ? wp[(int64_t)(j0 + q) * wstride] : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) tile[warp * 4 + q][lane] = pv[q];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jl = warp * 4 + q, j = j0 + jl;
            if (j < Hb && vok) {
                const float d = tile[lane][jl] + s_vec[j];
                const float db = s_vec[Hb + j];
                dp[(int64_t)(j0 + q) * dstride] = fmaf(gm, d, bt * db);
                ag = fmaf(wv[q], d, ag);
                ab = fmaf(wv[q], db, ab);
            }
        }
        __syncthreads();
    }
    s_ag[warp][lane] = ag; s_ab[warp][lane] = ab;
    __syncthreads();
    if (warp == 0 && vok) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += s_ag[k][lane]; b += s_ab[k][lane]; }
        dgamma[v] = a;
        dbeta[v] = b;
    }
}

}  // namespace tasu

using namespace tasu;

extern "C" int tasu_host_group_tokens(const int32_t* tok_host, int64_t n_rows, int V, int32_t* uniq_host,
                                      int32_t* seg_off_host, int32_t* perm_host, int32_t* n_uniq_host) {
    TASU_CHECK_ARG(n_rows >= 0 && V > 0, "n_rows >= 0, V > 0");
    TASU_CHECK_ARG(n_uniq_host != nullptr, "null n_uniq_host");
    *n_uniq_host = 0;
    if (n_rows == 0) { if (seg_off_host) seg_off_host[0] = 0; return TASU_OK; }
    TASU_CHECK_ARG(tok_host && uniq_host && seg_off_host && perm_host, "null pointer");
    thread_local std::vector<int32_t> count_tls;               // reused across calls (one per host thread)
    count_tls.assign((size_t)V, 0);
    int32_t* __restrict__ count = count_tls.data();            // hoisted: no TLS lookup inside the loops
    bool bad = false;
    for (int64_t r = 0; r < n_rows; ++r) {
        const uint32_t t = (uint32_t)tok_host[r];
        if (t >= (uint32_t)V) { bad = true; break; }
        ++count[t];
    }
    TASU_CHECK_ARG(!bad, "token id outside [0, V)");
    int32_t nu = 0, run = 0, sink_u = 0, sink_s = 0;
    for (int v = 0; v < V; ++v) {                               // branch-free: ~half of the bins are empty
        const int32_t c = count[v];
        int32_t* pu = c ? uniq_host + nu : &sink_u;
        int32_t* ps = c ? seg_off_host + nu : &sink_s;
        *pu = v; *ps = run;
        count[v] = run;                                        // exclusive start of token v
        nu += (c != 0); run += c;
    }
    seg_off_host[nu] = (int32_t)n_rows;
    for (int64_t r = 0; r < n_rows; ++r) perm_host[count[tok_host[r]]++] = (int32_t)r;    // stable: ascending r per token
    *n_uniq_host = nu;
    return TASU_OK;
}

// The whole host side of ctc_pseudo_posterior_noise (ps-slm.py:380-401, insert_prob = 0) in one call: u is ONE
// torch.rand(sum(L_b) + B) draw (the reference's stream: per utterance one uniform_ for alpha, then rand(L_b) for
// the keep mask); the descriptors go straight into the caller's (pinned) staging buffer, already grouped by token.
// Staging layout in int32 words, cap = sum(L_b):  uniq[cap] | seg_off[cap+1] | perm[cap] | hot[cap] | base[cap] |
// pad to 8 bytes | lens[B] (int64).
extern "C" int tasu_host_sim_token_rows(const float* u_host, const int32_t* tok_in_host, const int64_t* len_in_host, int B,
                                        int V, float drop_prob, float smooth_low, float smooth_high, int32_t* stage_host,
                                        int64_t stage_words, int64_t* n_rows_host, int32_t* n_uniq_host) {
    TASU_CHECK_ARG(B >= 0 && V > 0, "B >= 0, V > 0");
    TASU_CHECK_ARG(n_rows_host && n_uniq_host, "null output");
    *n_rows_host = 0; *n_uniq_host = 0;
    int64_t cap = 0;
    for (int b = 0; b < B; ++b) { TASU_CHECK_ARG(len_in_host[b] >= 0, "negative length"); cap += len_in_host[b]; }
    const int64_t o_uniq = 0, o_seg = cap, o_perm = 2 * cap + 1, o_hot = 3 * cap + 1, o_base = 4 * cap + 1;
    const int64_t o_lens = 5 * cap + 1 + ((5 * cap + 1) & 1);
    TASU_CHECK_ARG(stage_host && stage_words >= o_lens + 2 * (int64_t)B, "staging buffer too small");
    TASU_CHECK_ARG(B == 0 || (u_host && len_in_host && (cap == 0 || tok_in_host)), "null pointer");
    float* hot = reinterpret_cast<float*>(stage_host + o_hot);
    float* base = reinterpret_cast<float*>(stage_host + o_base);
    int64_t* lens = reinterpret_cast<int64_t*>(stage_host + o_lens);
    std::vector<int32_t> tok((size_t)cap);
    const volatile float span = smooth_high - smooth_low;       // fp32, like at::uniform_real_distribution<float>
    int64_t n = 0, iu = 0, it = 0;
    for (int b = 0; b < B; ++b) {
        const volatile float prod = u_host[iu++] * span;        // volatile: no fused multiply-add contraction
        const float alpha32 = prod + smooth_low;
        const double alpha = (double)alpha32;                   // .item() → python float
        const float a32 = (float)(1.0 - alpha), c32 = (float)(alpha / (double)V);
        const volatile float h32 = a32 + c32;
        int64_t kept = 0;
        for (int64_t i = 0; i < len_in_host[b]; ++i, ++iu, ++it) {
            if (u_host[iu] > drop_prob) {
                const int32_t t = tok_in_host[it];
                TASU_CHECK_ARG(t >= 0 && t < V, "token id outside [0, V)");
                tok[(size_t)n] = t; hot[n] = h32; base[n] = c32;
                ++n; ++kept;
            }
        }
        lens[b] = kept;
    }
    *n_rows_host = n;
    return tasu_host_group_tokens(tok.data(), n, V, stage_host + o_uniq, stage_host + o_seg, stage_host + o_perm, n_uniq_host);
}

extern "C" int tasu_linear_rowdots(const float* w1, int64_t w1_stride, const float* gamma, const float* beta,
                                   const float* b1, int N, int K, float* S, float* D, void* stream) {
    TASU_CHECK_ARG(N > 0 && K > 0 && w1_stride >= K, "shape");
    TASU_CHECK_ARG(w1 && gamma && beta && S && D, "null pointer");
    rowdots_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(w1, w1_stride, gamma, beta, b1, K, S, D);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_tokrow_fwd(const float* w1, int64_t w1_stride, const float* gamma, const float* S, const float* D,
                               const int32_t* uniq, const int32_t* seg_off, const int32_t* perm, const float* hot,
                               const float* base, int n_uniq, int64_t n_rows, int V, int Hb, float ln_eps, float* z,
                               void* h_bf16, float* row_a, float* row_e, const float* colT, void* stream) {
    TASU_CHECK_ARG(n_uniq >= 0 && n_rows >= 0 && V > 0 && Hb > 0 && w1_stride >= V, "shape");
    TASU_CHECK_ARG(Hb % 8 == 0, "Hb must be a multiple of 8");
    if (n_uniq == 0 || n_rows == 0) return TASU_OK;
    TASU_CHECK_ARG(w1 && gamma && S && D && uniq && seg_off && perm && hot && base && h_bf16 && row_a && row_e, "null pointer");
    TASU_CHECK_ARG(((uintptr_t)S % 16 == 0) && ((uintptr_t)D % 16 == 0) && ((uintptr_t)h_bf16 % 16 == 0) &&
                   (z == nullptr || (uintptr_t)z % 16 == 0) && (colT == nullptr || (uintptr_t)colT % 16 == 0), "16-byte alignment");
    const size_t smem = sizeof(float) * (size_t)Hb;
    TASU_CHECK_ARG(smem <= 200 * 1024, "Hb too large for the shared-memory column");
    if (smem > 48 * 1024)
        TASU_CHECK_CUDA(cudaFuncSetAttribute(tokrow_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tokrow_fwd_kernel, 256, smem) != cudaSuccess || per_sm <= 0) per_sm = 4;
    int64_t grid = (int64_t)sm_count() * per_sm;
    if (grid > n_uniq) grid = n_uniq;
    tokrow_fwd_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(w1, w1_stride, gamma, S, D, uniq, seg_off, perm, hot, base,
                                                                          n_uniq, V, Hb, ln_eps, z, (__nv_bfloat16*)h_bf16, row_a, row_e, colT);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_tokrow_rows_fwd(const float* colT, const float* S, const float* D, const int32_t* seg_off,
                                    const int32_t* perm, const float* hot, const float* base, int n_uniq, int64_t n_rows,
                                    int V, int Hb, float ln_eps, float* z, void* h_bf16, float* row_a, float* row_e,
                                    int32_t* row_slot_ws, void* stream) {
    TASU_CHECK_ARG(n_uniq >= 0 && n_rows >= 0 && V > 0 && Hb > 0 && Hb % 8 == 0, "shape (Hb multiple of 8)");
    if (n_uniq == 0 || n_rows == 0) return TASU_OK;
    TASU_CHECK_ARG(colT && S && D && seg_off && perm && hot && base && h_bf16 && row_a && row_e && row_slot_ws, "null pointer");
    TASU_CHECK_ARG(((uintptr_t)colT % 16 == 0) && ((uintptr_t)S % 16 == 0) && ((uintptr_t)D % 16 == 0) &&
                   ((uintptr_t)h_bf16 % 16 == 0) && (z == nullptr || (uintptr_t)z % 16 == 0), "16-byte alignment");
    cudaStream_t st = (cudaStream_t)stream;
    row_slot_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(seg_off, perm, n_uniq, n_rows, row_slot_ws);
    int64_t grid = (n_rows + 7) / 8, gmax = (int64_t)sm_count() * 8;
    if (grid > gmax) grid = gmax;
    tokrow_rows_kernel<<<(unsigned)grid, 256, 0, st>>>(colT, row_slot_ws, S, D, hot, base, n_rows, V, Hb, ln_eps, z,
                                                      (__nv_bfloat16*)h_bf16, row_a, row_e);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int64_t tasu_tokrow_cols_workspace(int V, int Hb) {
    if (V <= 0 || Hb <= 0) return 0;
    return (int64_t)((V + 31) / 32) * 2 * Hb * (int64_t)sizeof(float) + 4LL * ((V + 63) / 64 * 64);
}

extern "C" int tasu_tokrow_cols(const float* w1, int64_t w1_stride, const float* gamma, const float* beta, const float* b1,
                                const int32_t* uniq, int n_uniq, int V, int Hb, float* colT, float* S, float* D,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    TASU_CHECK_ARG(V > 0 && Hb > 0 && n_uniq >= 0 && w1_stride >= V, "shape");
    TASU_CHECK_ARG(w1 && gamma && beta && S && D && workspace, "null pointer");
    TASU_CHECK_ARG(n_uniq == 0 || (uniq && colT), "null uniq / colT");
    TASU_CHECK_ARG(workspace_bytes >= tasu_tokrow_cols_workspace(V, Hb), "workspace too small (tasu_tokrow_cols_workspace)");
    TASU_CHECK_ARG((uintptr_t)workspace % 16 == 0, "workspace alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_cta = (V + 31) / 32;
    float* part = (float*)workspace;
    int32_t* slot = (int32_t*)((char*)workspace + (int64_t)n_cta * 2 * Hb * sizeof(float));
    TASU_CHECK_CUDA(cudaMemsetAsync(slot, 0xFF, sizeof(int32_t) * V, st));
    if (n_uniq > 0) slot_scatter_kernel<<<(n_uniq + 255) / 256, 256, 0, st>>>(uniq, n_uniq, slot);
    tokrow_cols_kernel<<<n_cta, 256, 0, st>>>(w1, w1_stride, gamma, beta, slot, Hb, V, colT, part);
    TASU_CHECK_LAUNCH();
    reduce_partials_kernel<<<(2 * Hb + 31) / 32, 1024, 0, st>>>(part, n_cta, Hb, S, D, b1);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

static unsigned bwd_rows_grid(int Hb, int n_uniq) {
    const int chunks = (Hb + 2047) / 2048;
    switch (chunks) {
        case 1: return persistent_grid(tokrow_bwd_rows_kernel<1>, 256, n_uniq);
        case 2: return persistent_grid(tokrow_bwd_rows_kernel<2>, 256, n_uniq);
        case 3: return persistent_grid(tokrow_bwd_rows_kernel<3>, 256, n_uniq);
        default: return persistent_grid(tokrow_bwd_rows_kernel<4>, 256, n_uniq);
    }
}

extern "C" int64_t tasu_tokrow_bwd_workspace(int Hb, int n_uniq) {
    if (Hb <= 0 || n_uniq <= 0) return 0;
    return (int64_t)bwd_rows_grid(Hb, n_uniq) * 2 * Hb * (int64_t)sizeof(float);
}

extern "C" int tasu_tokrow_bwd_rows(const float* dh, const float* z, int64_t n_rows, int Hb, const int32_t* seg_off,
                                    const int32_t* perm, const float* row_a, const float* row_e, int n_uniq, float* P,
                                    float* db1, float* E, void* workspace, int64_t workspace_bytes, void* stream) {
    TASU_CHECK_ARG(n_rows >= 0 && n_uniq >= 0 && Hb > 0 && Hb % 8 == 0, "shape (Hb multiple of 8)");
    TASU_CHECK_ARG(Hb <= 4 * 256 * 8, "Hb <= 8192");
    TASU_CHECK_ARG(db1 && E, "null db1 / E");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0 || n_uniq == 0) {
        TASU_CHECK_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * Hb, st));
        TASU_CHECK_CUDA(cudaMemsetAsync(E, 0, sizeof(float) * Hb, st));
        return TASU_OK;
    }
    TASU_CHECK_ARG(dh && z && seg_off && perm && row_a && row_e && P && workspace, "null pointer");
    TASU_CHECK_ARG(((uintptr_t)dh % 16 == 0) && ((uintptr_t)z % 16 == 0) && ((uintptr_t)P % 16 == 0) &&
                   ((uintptr_t)workspace % 16 == 0), "16-byte alignment");
    const unsigned grid = bwd_rows_grid(Hb, n_uniq);
    TASU_CHECK_ARG(workspace_bytes >= (int64_t)grid * 2 * Hb * (int64_t)sizeof(float), "workspace too small (tasu_tokrow_bwd_workspace)");
    const int chunks = (Hb + 2047) / 2048;
    float* part = (float*)workspace;
#define LAUNCH(C) tokrow_bwd_rows_kernel<C><<<grid, 256, 0, st>>>(dh, z, Hb, seg_off, perm, row_a, row_e, n_uniq, P, part)
    if (chunks == 1) LAUNCH(1); else if (chunks == 2) LAUNCH(2); else if (chunks == 3) LAUNCH(3); else LAUNCH(4);
#undef LAUNCH
    TASU_CHECK_LAUNCH();
    reduce_partials_kernel<<<(2 * Hb + 31) / 32, 1024, 0, st>>>(part, (int)grid, Hb, db1, E, nullptr);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

extern "C" int tasu_tokrow_wgrad_finish(const float* P, const int32_t* uniq, int n_uniq, int32_t* slot_ws, const float* w1,
                                        int64_t w1_stride, const float* gamma, const float* beta, const float* E,
                                        const float* db1, int Hb, int V, float* dw1, int64_t dw1_stride, float* dgamma,
                                        float* dbeta, void* stream) {
    TASU_CHECK_ARG(Hb > 0 && V > 0 && n_uniq >= 0 && w1_stride >= V && dw1_stride >= V, "shape");
    TASU_CHECK_ARG(slot_ws && w1 && gamma && beta && E && db1 && dw1 && dgamma && dbeta, "null pointer");
    TASU_CHECK_ARG(n_uniq == 0 || (P && uniq), "null P / uniq");
    cudaStream_t st = (cudaStream_t)stream;
    TASU_CHECK_CUDA(cudaMemsetAsync(slot_ws, 0xFF, sizeof(int32_t) * V, st));
    if (n_uniq > 0) slot_scatter_kernel<<<(n_uniq + 255) / 256, 256, 0, st>>>(uniq, n_uniq, slot_ws);
    const size_t smem = 2 * sizeof(float) * (size_t)Hb;
    TASU_CHECK_ARG(smem <= 96 * 1024, "Hb too large");
    if (smem > 32 * 1024)
        TASU_CHECK_CUDA(cudaFuncSetAttribute(tokrow_wgrad_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tokrow_wgrad_finish_kernel<<<(unsigned)((V + 31) / 32), 256, smem, st>>>(P, slot_ws, w1, w1_stride, gamma, beta, E, db1, Hb, V,
                                                                            dw1, dw1_stride, dgamma, dbeta);
    TASU_CHECK_LAUNCH();
    return TASU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Composite entry points: the whole projector forward / backward of a text-only step behind ONE call each, so the
// host (Python) issues two calls instead of ~25 and the launches go out back to back.
static inline int64_t al256(int64_t x) { return (x + 255) / 256 * 256; }
static inline int64_t pad_to(int64_t n, int64_t a) { return (n + a - 1) / a * a; }

struct TrainWs {
    int64_t S, D, w2b, dyb, dh, P, part, E, slot, total;
};
static TrainWs train_ws(int64_t n, int n_uniq, int V, int Hb, int H) {
    TrainWs w;
    int64_t o = 0;
    w.S = o; o += al256(4LL * Hb);
    w.D = o; o += al256(4LL * Hb);
    w.w2b = o; o += al256(2LL * H * pad_to(Hb, 64));
    w.dyb = o; o += al256(2LL * (n > 0 ? n : 1) * pad_to(H, 8));
    w.dh = o; o += al256(4LL * (n > 0 ? n : 1) * Hb);
    w.P = o; o += al256(4LL * (n_uniq > 0 ? n_uniq : 1) * Hb);
    {
        const int64_t a = tasu_tokrow_bwd_workspace(Hb, n_uniq) + 16, b = tasu_tokrow_cols_workspace(V, Hb) + 16;
        w.part = o; o += al256(a > b ? a : b);
    }
    w.E = o; o += al256(4LL * Hb);
    w.slot = o; o += al256(4LL * (V > n ? V : n));      // finish: slot[V]; forward rows: row_slot[n_rows]
    w.total = o;
    return w;
}

extern "C" int64_t tasu_tokrow_train_workspace(int64_t n_rows, int n_uniq, int V, int Hb, int H) {
    if (n_rows < 0 || n_uniq < 0 || V <= 0 || Hb <= 0 || H <= 0) return 0;
    return train_ws(n_rows, n_uniq, V, Hb, H).total;
}

#define TASU_TRY(expr) do { int rc__ = (expr); if (rc__ != TASU_OK) return rc__; } while (0)

extern "C" int tasu_tokrow_linear_silu_fwd(const float* w1, int64_t w1_stride, const float* gamma, const float* beta,
                                           const float* b1, const float* w2, int64_t w2_stride, const float* b2,
                                           const int32_t* uniq, const int32_t* seg_off, const int32_t* perm,
                                           const float* hot, const float* base, int n_uniq, int64_t n_rows, int V, int Hb,
                                           int H, float ln_eps, float* z, void* h_bf16, float* row_a, float* row_e,
                                           void* y, int y_dtype, int64_t ldy, void* workspace, int64_t workspace_bytes,
                                           void* stream) {
    TASU_CHECK_ARG(n_rows >= 0 && n_uniq >= 0 && V > 0 && Hb > 0 && H > 0, "shape");
    if (n_rows == 0) return TASU_OK;
    TASU_CHECK_ARG(workspace && ((uintptr_t)workspace % 256 == 0), "workspace must be 256-byte aligned");
    const TrainWs ws = train_ws(n_rows, n_uniq, V, Hb, H);
    TASU_CHECK_ARG(workspace_bytes >= ws.total, "workspace too small (tasu_tokrow_train_workspace)");
    TASU_CHECK_ARG(w2 && b2 && y, "null pointer");
    char* wsb = (char*)workspace;
    float* S = (float*)(wsb + ws.S);
    float* D = (float*)(wsb + ws.D);
    void* w2b = wsb + ws.w2b;
    const int64_t ldw2 = pad_to(Hb, 64);
    float* colT = (float*)(wsb + ws.P);                      // the forward's compact columns share the backward's P area
    TASU_TRY(tasu_tokrow_cols(w1, w1_stride, gamma, beta, b1, uniq, n_uniq, V, Hb, colT, S, D, wsb + ws.part, ws.E - ws.part, stream));
    TASU_TRY(tasu_tokrow_rows_fwd(colT, S, D, seg_off, perm, hot, base, n_uniq, n_rows, V, Hb, ln_eps, z, h_bf16, row_a, row_e,
                                  (int32_t*)(wsb + ws.slot), stream));
    TASU_TRY(tasu_cast_rows(w2, TASU_F32, H, Hb, w2_stride, w2b, TASU_BF16, ldw2, nullptr, nullptr, 0.f, stream));
    TASU_TRY(tasu_gemm_bf16_tn(h_bf16, Hb, w2b, ldw2, y, y_dtype, ldy, (int)n_rows, H, Hb, TASU_EPI_BIAS, b2, nullptr, nullptr,
                               nullptr, nullptr, stream));
    return TASU_OK;
}

extern "C" int tasu_tokrow_linear_silu_bwd(const void* dy, int dy_dtype, int64_t ldy, const float* z, const void* h_bf16,
                                           const float* row_a, const float* row_e, const float* w1, int64_t w1_stride,
                                           const float* gamma, const float* beta, const float* w2, int64_t w2_stride,
                                           const int32_t* uniq, const int32_t* seg_off, const int32_t* perm, int n_uniq,
                                           int64_t n_rows, int V, int Hb, int H, float* dw1, int64_t dw1_stride,
                                           float* dgamma, float* dbeta, float* db1, float* dw2, int64_t dw2_stride,
                                           float* db2, int phase, void* workspace, int64_t workspace_bytes, void* stream) {
    TASU_CHECK_ARG(n_rows >= 0 && n_uniq >= 0 && V > 0 && Hb > 0 && H > 0, "shape");
    TASU_CHECK_ARG(phase >= 0 && phase <= 2, "phase: 0 = all, 1 = W1 half (dW1, dgamma, dbeta, db1), 2 = W2 half (dW2, db2)");
    TASU_CHECK_ARG(dw1 && dgamma && dbeta && db1 && dw2 && db2, "null gradient pointer");
    const bool do_w1 = phase != 2, do_w2 = phase != 1;
    TASU_CHECK_ARG(dy_dtype == TASU_F32 || dy_dtype == TASU_BF16, "dy_dtype");
    TASU_CHECK_ARG(workspace && ((uintptr_t)workspace % 256 == 0), "workspace must be 256-byte aligned");
    const TrainWs ws = train_ws(n_rows, n_uniq, V, Hb, H);
    TASU_CHECK_ARG(workspace_bytes >= ws.total, "workspace too small (tasu_tokrow_train_workspace)");
    cudaStream_t st = (cudaStream_t)stream;
    char* wsb = (char*)workspace;
    if (n_rows == 0) {                                      // no rows: every gradient is zero
        if (do_w1) {
            TASU_CHECK_CUDA(cudaMemsetAsync(db1, 0, 4LL * Hb, st));
            TASU_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, 4LL * V, st));
            TASU_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, 4LL * V, st));
            TASU_CHECK_CUDA(cudaMemset2DAsync(dw1, 4 * dw1_stride, 0, 4LL * V, Hb, st));
        }
        if (do_w2) {
            TASU_CHECK_CUDA(cudaMemsetAsync(db2, 0, 4LL * H, st));
            TASU_CHECK_CUDA(cudaMemset2DAsync(dw2, 4 * dw2_stride, 0, 4LL * Hb, H, st));
        }
        return TASU_OK;
    }
    const int64_t ldh8 = pad_to(H, 8), ldw2 = pad_to(Hb, 64);
    void* dyb = wsb + ws.dyb;
    void* w2b = wsb + ws.w2b;
    float* dh = (float*)(wsb + ws.dh);
    float* P = (float*)(wsb + ws.P);
    float* E = (float*)(wsb + ws.E);
    // The W1 half goes first: dW1 is 94 % of the gradient bytes, so a data-parallel caller can start its all-reduce
    // (phase 1 → all-reduce → phase 2) while the W2 half still computes.  The bf16 copy of dy made in phase 1 stays in
    // the workspace for phase 2 (nothing else may use the workspace in between).
    const void* dy_bf16 = dy;
    int64_t ld_dyb = ldy;
    if (dy_dtype != TASU_BF16 || (ldy * 2) % 16 != 0 || (uintptr_t)dy % 16 != 0) {
        if (do_w1) TASU_TRY(tasu_cast_rows(dy, dy_dtype, n_rows, H, ldy, dyb, TASU_BF16, ldh8, nullptr, nullptr, 0.f, stream));
        dy_bf16 = dyb; ld_dyb = ldh8;
    }
    // both contractions read dy, h and W2 where they lie (MN-major operands): no transposed copies
    if (do_w1) {
        TASU_TRY(tasu_cast_rows(w2, TASU_F32, H, Hb, w2_stride, w2b, TASU_BF16, ldw2, nullptr, nullptr, 0.f, stream));
        TASU_TRY(tasu_gemm_bf16_f32(dy_bf16, ld_dyb, 0, w2b, ldw2, 1, dh, Hb, (int)n_rows, Hb, H, stream));          // dh = dy · W2
        TASU_TRY(tasu_tokrow_bwd_rows(dh, z, n_rows, Hb, seg_off, perm, row_a, row_e, n_uniq, P, db1, E, wsb + ws.part,
                                      ws.E - ws.part, stream));
        TASU_TRY(tasu_tokrow_wgrad_finish(P, uniq, n_uniq, (int32_t*)(wsb + ws.slot), w1, w1_stride, gamma, beta, E, db1, Hb, V,
                                          dw1, dw1_stride, dgamma, dbeta, stream));
    }
    if (do_w2) {
        TASU_TRY(tasu_colsum(dy, dy_dtype, n_rows, H, ldy, db2, stream));
        TASU_TRY(tasu_gemm_bf16_f32(dy_bf16, ld_dyb, 1, h_bf16, Hb, 1, dw2, dw2_stride, H, Hb, (int)n_rows, stream));   // dW2 = dy^T · h
    }
    return TASU_OK;
}
