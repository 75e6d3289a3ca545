// Shared host/device structures of the persistent generation kernel.
#pragma once
#include <stdint.h>

#define SWN_NTHREADS 512
#define SWN_NWARPS (SWN_NTHREADS / 32)
#define SWN_MAX_SMEM 232448          // 227 KB opt-in dynamic shared memory per CTA on sm_100

enum { ROLE_LAYER = 0, ROLE_HEAD = 1, ROLE_IDLE = 2 };

// One descriptor per CTA of the persistent grid (global memory, read once at kernel start).
// All *_off fields are float offsets into GenParams::wbuf; sm_* are float offsets into the
// dynamic shared memory of the CTA.
struct CtaDesc {
    int32_t role;
    int32_t layer;        // layer index (ROLE_LAYER) or head CTA index inside its head group (ROLE_HEAD)
    int32_t hgroup;       // head group (ROLE_HEAD): owns the pipeline slots m with m % NHG == hgroup
    int32_t dil;          // dilation of dil_h[layer]
    int32_t pair0, npairs;      // gate pairs (z row i, candidate row H+i) owned: i in [pair0, pair0+npairs)
    int32_t skip0, nskip;       // rows of out_skip[layer-1] owned (layer >= 1)
    int32_t smem_taps;          // taps K-smem_taps .. K-1 of dil_h are SMEM resident, the rest is read from wbuf
    int32_t skip_in_smem;
    int32_t ring_in_smem;
    int32_t ring_len;           // RL = (K-1)*dil + SEG
    int32_t sm_w, sm_ws, sm_h, sm_ring;
    int64_t w_off;              // [K][npairs][2][Hp]
    int64_t b_off;              // [npairs][2]
    int64_t ws_off;             // [nskip][Hp]
    int64_t ring_off;           // float offset into ring_scratch: [M][RL][npairs][2][BS]
    // head CTA: row ranges of the three head mat-vecs and of the input stage
    int32_t a0, na;             // rows of sum-of-skips (S)
    int32_t b0, nb;             // rows of out_1 (n1)
    int32_t c0, nc;             // rows of out_2 (n2)  (CSWNV: all rows in every head CTA)
    int32_t i0, ni;             // rows of the input stage (H)
    int32_t wa_in_smem, wb_in_smem, wc_in_smem;
    int32_t sm_wa, sm_wb, sm_wc, sm_x0, sm_x1, sm_x2, sm_misc;
    int64_t wa_off, ca_off;     // out_skip[L-1] rows [na][Hp], constants [na] (sum of all skip biases)
    int64_t wb_off, bb_off;     // out_1 rows [nb][Sp], bias [nb]
    int64_t wc_off, bc_off;     // out_2 rows [nc][n1p], bias [nc]
};

struct GenParams {
    int32_t kind, H, Hp, S, Sp, K, L, lpc, n1, n1p, n2, Q, U, R, B, M, Tf;
    int32_t t_first, it0, n_it, n_iter;
    int32_t NH, NHG, audio_in;
    int64_t n_out;                  // generated samples per utterance = n_iter * SEG
    const CtaDesc* desc;
    const float* wbuf;
    // conditioning
    const float* P;                 // [B][Tf][L][SEG][2H]
    const float* pconst;            // [L][2H]
    const float* gain;              // [U]
    // input stage
    const float* in_tab;            // CSWNV: [K][2][H] (cb_j, v_j); DSWNV: [K][Q][H]
    const float* in_bias;           // [H]
    const float* xa_tab;            // DSWNV audio_in_flag: [L][Q][2H] extra in_x columns, else null
    // LL exchange buffers (64-bit words: low = fp32 bits, high = tag)
    unsigned long long* ring_ll;    // [L+1][M][NC][H]
    unsigned long long* skip_ll;    // [L-1][M][BS][S]
    unsigned long long* hx1;        // [M][BS][S]
    unsigned long long* hx2;        // [M][BS][n1]
    unsigned long long* hx3;        // [M][NH][BS][2]
    unsigned long long* idx_ll;     // [M][BS]   sampled / fed-back index of the step (audio_in_flag)
    float* ring_scratch;
    // io
    const float* noise;
    unsigned long long seed;
    const void* forced;
    const void* seedw;              // optional seed waveform [B][SEG]
    void* out;
    float* head_out;
    int32_t* status;                // [0] abort flag, [1] where, [2] iteration
    unsigned long long timeout_cycles;
    unsigned long long* prof;       // optional [prof_count][grid][32] globaltimer stamps
    int32_t prof_start, prof_count;
    int32_t poll_mode, poll_sleep_ns;
};
