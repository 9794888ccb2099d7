/* Plain-C client of libtasu_bridge.so: no torch, no C++ — only the C ABI of include/tasu_bridge.h and the CUDA runtime
 * for device memory.  Runs the PSD decision table of SURVEY.md §8(a) (ps-slm.py:265-297): greedy ids
 * [0,0,5,5,5,0,7,7,0,5,0,0] with blank probabilities [.95,.6,.1,.2,.3,.91,.05,.05,.89,0,.9,.8999] must keep 6
 * candidates — starts [1,2,6,8,9,11], lengths [1,3,2,1,1,1] (blank frames below 0.9 individually, 0.9 itself dropped).
 * Build: gcc abi_smoke.c -I../../include -I/usr/local/cuda/include -L<libdir> -ltasu_bridge -L/usr/local/cuda/lib64 -lcudart */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tasu_bridge.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e_)); return 2; } } while (0)
#define TK(x) do { int r_ = (x); if (r_ != TASU_OK) { printf("tasu error %d: %s\n", r_, tasu_last_error()); return 3; } } while (0)

int main(void) {
    enum { B = 1, T = 12, V = 9 };
    const int ids[T] = {0, 0, 5, 5, 5, 0, 7, 7, 0, 5, 0, 0};
    const float pb[T] = {.95f, .6f, .1f, .2f, .3f, .91f, .05f, .05f, .89f, 0.f, .9f, .8999f};
    float post[T * V];
    for (int t = 0; t < T; ++t) {
        /* probability rows: blank gets pb, the greedy id gets the largest share of the rest (or blank stays the max) */
        const float rest = 1.f - pb[t];
        for (int v = 0; v < V; ++v) post[t * V + v] = 0.f;
        post[t * V + 0] = pb[t];
        if (ids[t] != 0) post[t * V + ids[t]] = rest;
        else { for (int v = 1; v < V; ++v) post[t * V + v] = rest / (V - 1); }
    }
    int sm = 0, major = 0, minor = 0;
    TK(tasu_device_info(&sm, &major, &minor));
    printf("abi %d, device sm_%d%d with %d SMs\n", tasu_abi_version(), major, minor, sm);

    float *d_post, *d_xb, *d_mx, *d_score; int32_t *d_arg, *d_ss, *d_sl, *d_kf, *d_sfo, *d_roff, *d_foff, *d_cnt;
    uint32_t* d_gmax; int64_t *d_lens, *d_new, *d_hdr;
    const int64_t lens = T;
    CK(cudaMalloc((void**)&d_post, sizeof(post))); CK(cudaMemcpy(d_post, post, sizeof(post), cudaMemcpyHostToDevice));
    CK(cudaMalloc((void**)&d_xb, 4 * T)); CK(cudaMalloc((void**)&d_mx, 4 * T)); CK(cudaMalloc((void**)&d_score, 4 * T));
    CK(cudaMalloc((void**)&d_arg, 4 * T)); CK(cudaMalloc((void**)&d_ss, 4 * T)); CK(cudaMalloc((void**)&d_sl, 4 * T));
    CK(cudaMalloc((void**)&d_kf, 4 * B)); CK(cudaMalloc((void**)&d_sfo, 4 * T)); CK(cudaMalloc((void**)&d_roff, 4 * (B + 1)));
    CK(cudaMalloc((void**)&d_foff, 4 * (B + 1))); CK(cudaMalloc((void**)&d_cnt, 16)); CK(cudaMalloc((void**)&d_gmax, 4));
    CK(cudaMalloc((void**)&d_lens, 8)); CK(cudaMemcpy(d_lens, &lens, 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc((void**)&d_new, 8 * B)); CK(cudaMalloc((void**)&d_hdr, 8 * TASU_CH_WORDS));

    TK(tasu_frame_stats(d_post, TASU_F32, TASU_INPUT_PROBS, B, T, V, (int64_t)T * V, V, 0, NULL, d_arg, d_xb, d_mx, NULL, d_gmax, NULL));
    TK(tasu_collapse_plan(d_arg, d_xb, d_mx, NULL, d_gmax, TASU_INPUT_PROBS, d_lens, B, T, 0, 0.90f, d_ss, d_sl, d_score, d_new,
                          d_kf, d_sfo, NULL));
    TK(tasu_collapse_scan(d_new, d_kf, d_gmax, B, d_roff, d_foff, d_hdr, d_cnt, NULL));
    CK(cudaDeviceSynchronize());

    int32_t arg[T], ss[T], sl[T]; float score[T]; int64_t hdr[TASU_CH_WORDS], m = 0;
    CK(cudaMemcpy(arg, d_arg, sizeof(arg), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ss, d_ss, sizeof(ss), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sl, d_sl, sizeof(sl), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(score, d_score, sizeof(score), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdr, d_hdr, sizeof(hdr), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&m, d_new, 8, cudaMemcpyDeviceToHost));
    const int want_start[6] = {1, 2, 6, 8, 9, 11}, want_len[6] = {1, 3, 2, 1, 1, 1};
    const float want_score[6] = {.6f, .2f, .05f, .89f, 0.f, .8999f};
    int ok = (m == 6) && hdr[TASU_CH_N_OUT] == 6 && hdr[TASU_CH_MAX_LEN] == 6 && hdr[TASU_CH_IS_LOGPROB] == 0;
    for (int t = 0; t < T; ++t) ok = ok && arg[t] == ids[t];
    for (int j = 0; j < 6 && ok; ++j) {
        float d = score[j] - want_score[j];
        ok = ss[j] == want_start[j] && sl[j] == want_len[j] && d < 1e-6f && d > -1e-6f;
    }
    /* grouped kept-frame layout (step 2b'): run lengths {1,3,2,1,1,1} -> 4 single rows, one 2-frame and one 4-row group */
    {
        int32_t *d_slot, *d_xoff, *d_cntb, *d_lay, *d_ticket;
        CK(cudaMalloc((void**)&d_slot, 4 * B * T)); CK(cudaMalloc((void**)&d_xoff, 4 * B * T));
        CK(cudaMalloc((void**)&d_cntb, 4 * 8 * B)); CK(cudaMalloc((void**)&d_lay, 4 * TASU_GL_WORDS));
        CK(cudaMalloc((void**)&d_ticket, 4)); CK(cudaMemset(d_ticket, 0, 4));
        TK(tasu_group_plan(d_sl, d_new, B, T, d_slot, d_xoff, d_cntb, d_cntb + 4 * B, d_lay, d_ticket, NULL));
        CK(cudaDeviceSynchronize());
        int32_t lay[TASU_GL_WORDS], slot[T];
        CK(cudaMemcpy(lay, d_lay, sizeof(lay), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(slot, d_slot, sizeof(slot), cudaMemcpyDeviceToHost));
        const int want_slot[6] = {0, 0, 0, 1, 2, 3};
        ok = ok && lay[TASU_GL_NS] == 4 && lay[TASU_GL_N2] == 1 && lay[TASU_GL_N4] == 1 && lay[TASU_GL_NXE] == 0 &&
             lay[TASU_GL_A2] == 128 && lay[TASU_GL_A4] == 256 && lay[TASU_GL_AX] == 384 && lay[TASU_GL_A_ROWS] == 384 &&
             lay[TASU_GL_O4] == 192 && lay[TASU_GL_OX] == 224;
        for (int j = 0; j < 6; ++j) ok = ok && slot[j] == want_slot[j];
    }
    /* invalid arguments are reported, never thrown */
    ok = ok && tasu_frame_stats(NULL, TASU_F32, TASU_INPUT_PROBS, 1, 1, 0, 0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL) == TASU_ERR_INVALID_ARG;
    printf("kept %lld candidates: %s\n", (long long)m, ok ? "C ABI SMOKE OK" : "MISMATCH");
    return ok ? 0 : 1;
}
