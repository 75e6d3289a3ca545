// Pipelined batch generation kernel (variant 3) of the B200-native shallow-wavenet path.
//
// Same algorithm as swn_gen.cuh (CSWNV.batch_fast_generate, src/nets/cswnv_shift1.py:314-402;
// DSWNV.batch_fast_generate, src/nets/dswnv.py:312-374) but organised for THROUGHPUT over many
// utterances: the batch is cut into M pipeline slots of BS utterances (NC = SEG*BS columns);
// the L layer groups and the head CTAs form a ring pipeline, every stage working on a different
// slot at any time.  Differences to variant 1:
//   * a layer CTA evaluates all taps of its gate pairs (and its out_skip rows) as ONE
//     register-tiled fp32 GEMM  [rows x H] . [H x NC]  against weights that stay in shared memory
//     (k-major, so a thread reads its TR rows with 128-bit loads) - no shuffle reductions;
//   * the CTA is warp-specialised: GEMM warps, one loader lane that waits for the upstream
//     counter and pulls the next slot's input tile with one TMA bulk copy (double buffered,
//     mbarrier completion), and epilogue warps that prefetch conditioning + ring partial sums,
//     apply the gate, publish, and scatter the delayed taps into the ring off the critical path;
//   * tiles travel through L2 as plain fp32 [H][NC] blocks plus one monotonic counter per
//     (stage, slot) (release add / acquire poll) instead of {value,tag} words;
//   * out_skip of the last layer is evaluated by the last layer's CTAs on the rows they own
//     (partial sums over k, added by the head in a fixed order); every head CTA serves whole
//     slots on its own, streaming out_1 / out_2 through a TMA ring.
// Column order inside a slot: col = s*BS + b (time-major), so that the BS utterances of one time
// column are contiguous in tiles and rings.
#pragma once
#include "swn_gen.cuh"

#ifndef B3_ABL
#define B3_ABL 0     // diagnostic builds only (tools/run_r02_v3abl.sh): 1 = no ring scatter, 2 = no tile stores of the gated cells, 3 = no tile stores of the tail; results are wrong, DRAM traffic is what is measured
#endif

namespace swn {

constexpr int B3_NGW = 4;                 // GEMM warps 0..3 (one per SM sub-partition)
constexpr int B3_LOADER = 4;              // warp 4: lane 0 is the loader
constexpr int B3_EPI0 = 5;                // epilogue warps 5..15
constexpr int B3_CNT_STRIDE = 8;          // one counter per 32-byte sector
constexpr int B3_HIST = 32;
// Packed fp32 FMA (FFMA2, sm_100) in the GEMM inner loop.  Measured on B200: 7.4 us per 128x192x40 tile against
// 6.8 us with scalar FFMA (the packed form halves the issue slots but not the FMA-pipe cycles), so it stays off.
#ifndef B3_TMA_SPLIT
#define B3_TMA_SPLIT 1
#endif
#ifndef B3_WAIT_SLEEP_NS
#define B3_WAIT_SLEEP_NS 64
#endif
constexpr int B3_P1W = 3;                 // warps of the gate-operand prefetch team (p1_split)
constexpr int B3_NBW = 4;                 // cell 0 with build_x: warps 5..8 build the input tile, 9..15 are the epilogue

enum { ROLE_DENSE = 3 };                  // besides ROLE_LAYER (gated cell), ROLE_HEAD (tail), ROLE_IDLE

// One descriptor per CTA.  Pipeline stages: the L gated cells, then the dense stages of the output head
// (A: relu(sum of skips) = out_skip[L-1] . h + the other layers' skips, B: relu(out_1), C: out_2 when it is
// large), then the tail CTAs (small out_2, sampler, causal input stage of the next step).
struct BCta {
    int32_t role, layer, hindex;          // layer: cell index / dense stage id; hindex: index inside the group / tail index
    int32_t dil;
    int32_t pair0, npairs, skip0, nskip;  // gated cell: owned gate pairs and rows of out_skip[layer-1]
    int32_t row0, nrows;                  // dense stage: owned output rows
    int32_t rows, RP;                     // GEMM rows and their padding to 8
    int32_t kdim, kp;                     // GEMM depth (input channels) and its padding to 4
    int32_t ring_len, ring_in_smem;
    int32_t cnt_in, cnt_out, nprod_in;    // counter rows ([.][M]) of the input / output tile; producers of the input
    int32_t nxb;                          // input buffers in shared memory (2, or 1 when two large tiles do not fit)
    int32_t is_last, relu, add_skips;
    int32_t in_slot_f, out_slot_f;        // floats per slot of the input / output tile
    int32_t sm_w, sm_x, sm_o, sm_pre, sm_ring, sm_bar, sm_build;
    int64_t in_off, out_off;              // float offsets of slot 0 inside BatchParams::tiles
    int64_t w_off;                        // [kp][RP]
    int64_t wr_off;                       // umma: the same rows row-major [128][kp] (zero padded), loaded into tensor memory
    int64_t b_off;                        // gated cell: [np][2]; dense: [nrows]
    int64_t ring_off;                     // [RL][2*np][Bp] in ring_scratch
    // tail CTA shared-memory layout
    int32_t sm_x2, sm_oc, sm_w2, sm_tin, sm_itab, sm_ain, sm_noise, sm_hist, sm_lpb, sm_misc;
    int32_t tail_stride;                  // floats of one tail instance's shared-memory block
};

struct BatchParams {
    int32_t kind, H, S, K, L, lpc, n1, n1p, n2, n2p, Q, U, R, B, M, Tf;
    int32_t t_first, it0, n_it, n_iter;
    int32_t NT;                           // tail CTAs
    int32_t tail_inst;                    // tail instances per tail CTA (1 x 512 threads or 2 x 256)
    int32_t p1_split;                     // gated cells: a separate 3-warp team prefetches the gate operands one pass ahead
    int32_t tail_gemm;                    // 1: the tail evaluates out_2 itself (few rows); 0: out_2 is dense stage C
    int32_t tail_cnt, tail_nprod, tail_slot_f;   // counter row / producers / floats per slot of the tail's input
    int64_t tail_in_off;
    int64_t n_out;
    const BCta* desc;
    const float* wbuf;
    const float* P;                       // [B][Tf][L][SEG][2H]
    const float* pconst;                  // [L][2H]
    const float* gain;                    // [U]
    const float* in_tab;
    const float* in_bias;
    int64_t w2_off, b2_off;               // tail_gemm: out_2 row-major [n2][n1p] and its bias, inside wbuf
    int64_t cs_off;                       // sum over layers of the out_skip biases [S]
    float* tiles;                         // all exchange tiles (see the host plan)
    unsigned int* cnt;                    // [stages+1][M] * B3_CNT_STRIDE
    float* skipbuf;                       // [L-1][M][S][BS]
    float* hist_g;                        // build_x: [M][BS][HIST] fed-back samples (written by the tail, read by cell 0)
    int32_t build_x;                      // 1: the CTAs of cell 0 evaluate the causal input stage themselves
    const int32_t* slot_nit;              // [M] loop steps slot m still has to generate (utterances of unequal length)
    int32_t ndb;                          // umma: accumulator tiles in tensor memory (2, or 1 when 2H + 2NP > 512 columns)
    int32_t umma, NP;                     // 1: the gated cells contract on the tensor cores (tcgen05, 3 x tf32); NP = columns padded to 16
    float* ring_scratch;
    const float* noise;
    unsigned long long seed;
    const void* forced;
    const void* seedw;
    void* out;
    float* head_out;
    int32_t* status;
    unsigned long long timeout_cycles;
    unsigned long long* prof;             // optional [prof_count][M][grid][16]
    int32_t prof_start, prof_count;
};

// ------------------------------------------------------------------ primitives
__device__ __forceinline__ uint32_t b3_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void b3_mbar_init(unsigned long long* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b3_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void b3_mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b3_smem(b)) : "memory");
}
__device__ __forceinline__ void b3_mbar_arm(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b3_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool b3_mbar_try(unsigned long long* b, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(b3_smem(b)), "r"(parity) : "memory");
    return done != 0;
}
// wait with an escape: returns false when another CTA / role aborted
__device__ __forceinline__ bool b3_mbar_wait(unsigned long long* b, uint32_t parity, const int32_t* status) {
    uint32_t spins = 0;
    while (!b3_mbar_try(b, parity)) {
#if B3_WAIT_SLEEP_NS
        __nanosleep(B3_WAIT_SLEEP_NS);       // waiting warps must not take issue slots from the working ones
#endif
        if (((++spins) & 63u) == 0 && ((volatile const int32_t*)status)[0] != 0) return false;
    }
    return true;
}
// warp-uniform variant: every lane polls (so each observes the completion itself), the verdict is shared
__device__ __forceinline__ bool b3_wait_warp(unsigned long long* b, uint32_t parity, const int32_t* status) {
    const bool ok = b3_mbar_wait(b, parity, status);
    return __all_sync(0xffffffffu, ok);
}
__device__ __forceinline__ void b3_tma_load(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(b3_smem(smem_dst)), "l"(gsrc), "r"(bytes), "r"(b3_smem(bar)) : "memory");
}
__device__ __forceinline__ void b3_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Spin until the monotonic counter reached `target` (acquire); false = watchdog / abort.
static __device__ __noinline__ bool b3_cnt_wait(const unsigned int* p, unsigned int target, const BatchParams& P, int where) {
    const long long t_start = clock64();
    uint32_t spins = 0;
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v >= target) return true;
        if (((++spins) & 63u) == 0) {
            volatile int32_t* st = P.status;
            if (st[0] != 0) return false;
            if ((unsigned long long)(clock64() - t_start) > P.timeout_cycles) {
                if (atomicCAS((int*)P.status, 0, 1) == 0) { st[1] = where; st[2] = (int)target; st[3] = blockIdx.x; }
                return false;
            }
        }
    }
}
// Release increment: the stores of the whole team (ordered before this thread by its barrier) become visible
// to whoever acquires the counter.
__device__ __forceinline__ void b3_signal(unsigned int* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// ---- tcgen05 pieces (validated stand-alone by tools/microbench/umma_tf32.cu)
// Tile layout of a cell input in UMMA mode: B operand [n][k], K-major, no swizzle: 128-byte core matrices of
// 8 columns x 4 k; float offset of (column n, channel k) inside one (hi or lo) tile of NP columns:
__device__ __forceinline__ int b3_uoff(int n, int k, int NP) { return (((k >> 2) * (NP >> 3) + (n >> 3)) << 5) + ((n & 7) << 2) + (k & 3); }
__device__ __forceinline__ uint64_t b3_udesc(uint32_t saddr, int NP) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((((NP >> 3) * 128) >> 4) & 0x3FFF) << 16;     // leading byte offset: next core matrix along k
    d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;                   // stride byte offset: next 8-column group
    d |= (uint64_t)1 << 46;                                       // descriptor version (Blackwell)
    return d;                                                     // layout type 0: no swizzle
}
__device__ __forceinline__ void b3_umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
// the same with the accumulate flag fixed to 1 (no predicate set-up in front of every issue)
__device__ __forceinline__ void b3_umma_tf32_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile("tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, 1;" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void b3_umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b3_smem(bar)) : "memory");
}
__device__ __forceinline__ void b3_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void b3_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ float b3_tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Passes of a stage, in ring order (iteration-major, slot-minor).  A slot whose utterances are all complete drops
// out of the ring: every role of every CTA skips the same passes, `q` counts the passes actually processed
// (buffer / mbarrier parities follow it), `it` and `m` are the loop iteration and the slot.
#define B3_FOR_PASSES(q, it, m)                                                        \
    for (int q##_raw = 0, q = -1, it = 0, m = 0; q##_raw < nq; ++q##_raw)               \
        if (it = q##_raw / M, m = q##_raw - it * M, (it - P.it0 < __ldg(P.slot_nit + m)) ? (++q, true) : false)

#define B3_PROF(ev)                                                                                      \
    do {                                                                                                 \
        if (P.prof && it >= P.prof_start && it < P.prof_start + P.prof_count)                            \
            P.prof[((((size_t)(it - P.prof_start) * P.M + m) * gridDim.x + blockIdx.x) << 4) + (ev)] = gtime(); \
    } while (0)

__host__ __device__ constexpr int b3_pick_tc(int nc) {
    return (nc % 5 == 0) ? 5 : (nc % 4 == 0) ? 4 : (nc % 3 == 0) ? 3 : (nc % 2 == 0) ? 2 : 1;
}

// ------------------------------------------------------------------ GEMM stage CTA (gated cell or dense stage)
// Register tile of the fp32 GEMM for NCX columns: a thread owns TR rows x TC columns, a warp NCG column groups x
// RGW row groups; KS warps split k for the same rows (their partial sums are added through the output tile).
// Shared-memory loads, not FFMA issue, bound small tiles (every LDS.128 costs 4 LSU cycles even when it is a
// broadcast), hence the large tiles where the column count allows them.
template <int NCX, bool DENSE = false>
struct B3Tile {
    static constexpr int TC = (NCX % 10 == 0 && NCX >= 40) ? 10 : (NCX <= 8) ? NCX : b3_pick_tc(NCX);
    static constexpr int NCG = NCX / TC;            // column groups (NCG*TC == NCX)
    static constexpr int RGW = 32 / NCG;            // row groups handled by one warp pass
    // dense head stages own 64..256 rows and few columns: 4-row tiles (2 when a warp pass would cover >= 128 rows)
    static constexpr int TR = DENSE ? ((32 / NCG) * 4 >= 128 ? 2 : 4) : (NCX == 20 || NCX >= 40) ? 8 : 4;
    static constexpr int RPW = RGW * TR;            // rows per warp pass
    static constexpr int KS = (RPW >= 64) ? 2 : 1;  // k-splits
    static constexpr int OS = NCX | 1;
    static_assert(NCG >= 1 && NCG <= 32 && NCG * TC == NCX, "column groups");
};

// MODE 0: gated cell (cswnv_shift1.py:275-285) over NC = SEG*BS columns; MODE 1: dense head stage over BS columns.
template <int SEG, int BS, int MODE>
__device__ void b3_stage_body(const BatchParams& P, const BCta& D, float* smem);

template <int SEG, int BS, int MODE>
__device__ void b3_stage(const BatchParams& P, const BCta& D, float* smem) {
    b3_stage_body<SEG, BS, MODE>(P, D, smem);            // every role returns here, also after an abort
    if (MODE == 0 && P.umma) {
        b3_tc_fence_before();
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t tb = *reinterpret_cast<const uint32_t*>(reinterpret_cast<unsigned long long*>(smem + D.sm_bar) + 14);
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
        }
    }
}

template <int SEG, int BS, int MODE>
__device__ void b3_stage_body(const BatchParams& P, const BCta& D, float* smem) {
    constexpr int NC = MODE == 0 ? SEG * BS : BS;
    using T = B3Tile<NC, MODE == 1>;
    constexpr int TC = T::TC, NCG = T::NCG, RGW = T::RGW, TR = T::TR, RPW = T::RPW, OS = T::OS, KS = T::KS;
    constexpr int SLOTS = B3_NGW / KS;              // row blocks in flight
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = D.layer, np = D.npairs, K = P.K, H = P.H, kp = D.kp, dil = D.dil, RL = D.ring_len, RP = D.RP;
    const int M = P.M, Bp = M * BS, np2 = 2 * np;

    float* Wt = smem + D.sm_w;
    float* Xs = smem + D.sm_x;
    float* Os = smem + D.sm_o;
    float4* pre = reinterpret_cast<float4*>(smem + D.sm_pre);
    float* pre1 = smem + D.sm_pre;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + D.sm_bar);
    unsigned long long* x_full = bars;              // [2]
    unsigned long long* x_empty = bars + 2;         // [2]
    unsigned long long* o_full = bars + 4;
    unsigned long long* o_empty = bars + 5;
    unsigned long long* ebar = bars + 6;            // barrier among the epilogue warps (abortable, unlike bar.sync)
    unsigned long long* kbar = bars + 7;            // [SLOTS] first k-split stored its partial sums
    unsigned long long* kback = bars + 7 + SLOTS;   // [SLOTS] second k-split added its own (only waited for between passes)

    const bool umma = MODE == 0 && P.umma;                    // tensor-core contraction (weights in tensor memory)
    // buffer index / use count of pass q for the input buffers (nxb of them) and the accumulator tiles (ndb)
    const int nxb = D.nxb, ndb = P.ndb;
    auto XB = [nxb](int q) { return nxb == 2 ? (q & 1) : 0; };
    auto XU = [nxb](int q) { return (uint32_t)(nxb == 2 ? (q >> 1) : q); };
    auto DB = [ndb](int q) { return ndb == 2 ? (q & 1) : 0; };
    auto DU = [ndb](int q) { return (uint32_t)(ndb == 2 ? (q >> 1) : q); };
    const int NP = P.NP, utile = D.kdim * NP;                 // umma: floats of one (hi or lo) input tile
    const bool build = MODE == 0 && P.build_x && l == 0 && !umma;   // this CTA evaluates the causal input stage itself
    const bool ubuild = umma && P.build_x && l == 0;          // same, tensor-core mode: done by the (mostly idle) drain warps
    // throughput-bound rings: three warps (p1w0 ..) prefetch the gate operands of the NEXT pass while the others gate this one
    const bool psplit = MODE == 0 && P.p1_split && M >= 2;
    const int p1w0 = build ? B3_EPI0 + B3_NBW : (umma ? B3_EPI0 + 1 : B3_EPI0);
    const int epi0 = p1w0 + (psplit ? B3_P1W : 0);            // first warp of the (main) epilogue team
    unsigned long long* pre_full = bars + 16;       // [2] psplit: operands of a pass prefetched
    unsigned long long* pre_empty = bars + 18;      // [2] psplit: ... consumed
    unsigned long long* d_full = bars + 7;          // umma: [2] accumulator tile complete (tcgen05.commit)
    unsigned long long* d_empty = bars + 9;         // umma: [2] accumulator tile drained
    unsigned long long* x_lo = bars + 11;           // umma: [2] low-half operand tile landed (the high half completes x_full)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    const int nepi = (SWN_NWARPS - epi0) * 32;
    unsigned long long* bbar = bars + 15;           // barrier among the builder warps

    const int nrb = (RP + RPW - 1) / RPW;
    const int nslot = nrb < SLOTS ? nrb : SLOTS;    // row-block slots in use
    const int ngwa = umma ? 4 : nslot * KS;         // GEMM warps that own work (umma: the four drain warps)

    // ---- one-time staging
    {
        if (!umma) {
            const float4* src = reinterpret_cast<const float4*>(P.wbuf + D.w_off);
            float4* dst = reinterpret_cast<float4*>(Wt);
            const int n4 = kp * RP / 4;
            for (int i = tid; i < n4; i += SWN_NTHREADS) dst[i] = src[i];
        }
        const int nx = umma ? nxb * 2 * utile : nxb * kp * NC;
        for (int i = tid; i < nx; i += SWN_NTHREADS) Xs[i] = 0.f;
        if (build || ubuild) {
            // in_tab [K][2][H], then cst[r] = in_bias[r] + sum_j cb_j[r] (all taps inside the sequence)
            float* itab = smem + D.sm_build;
            for (int i = tid; i < K * 2 * H; i += SWN_NTHREADS) itab[i] = P.in_tab[i];
            for (int r = tid; r < H; r += SWN_NTHREADS) {
                float c = P.in_bias[r];
                for (int j = 0; j < K; ++j) c += P.in_tab[(j * 2) * H + r];
                itab[K * 2 * H + r] = c;
                itab[K * 2 * H + H + r] = P.in_bias[r];
            }
        }
        if (MODE == 0 && D.ring_in_smem) {
            float* rg = smem + D.sm_ring;
            for (int i = tid; i < RL * np2 * Bp; i += SWN_NTHREADS) rg[i] = 0.f;
        }
        if (tid == 0) {
            b3_mbar_init(&x_full[0], (build || ubuild) ? B3_NBW : 1); b3_mbar_init(&x_full[1], (build || ubuild) ? B3_NBW : 1);
            b3_mbar_init(bbar, B3_NBW);
            b3_mbar_init(&x_empty[0], umma ? 2 : ngwa + 1); b3_mbar_init(&x_empty[1], umma ? 2 : ngwa + 1);
            b3_mbar_init(o_full, ngwa); b3_mbar_init(o_empty, 1);
            b3_mbar_init(ebar, SWN_NWARPS - epi0);
            b3_mbar_init(&pre_full[0], B3_P1W); b3_mbar_init(&pre_full[1], B3_P1W);
            b3_mbar_init(&pre_empty[0], 1); b3_mbar_init(&pre_empty[1], 1);
            if (umma) {
                b3_mbar_init(&x_lo[0], ubuild ? B3_NBW : 1); b3_mbar_init(&x_lo[1], ubuild ? B3_NBW : 1);
                b3_mbar_init(&d_full[0], 1); b3_mbar_init(&d_full[1], 1);
                b3_mbar_init(&d_empty[0], 4); b3_mbar_init(&d_empty[1], 4);
            } else {
                for (int i = 0; i < SLOTS; ++i) { b3_mbar_init(&kbar[i], 1); b3_mbar_init(&kback[i], 1); }
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (umma && warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(b3_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    b3_fence_proxy_async();         // the zeroed X buffers are overwritten by the async proxy later
    b3_tc_fence_before();
    __syncthreads();
    b3_tc_fence_after();
    const uint32_t tbase = umma ? *tmem_slot : 0u;
    if (umma && warp < 4) {
        // weights -> tensor memory: row = lane, channel k = column; tf32 high halves at [0,kp), low halves at [kp,2kp)
        const float* wr = P.wbuf + D.wr_off + (size_t)tid * kp;
        const uint32_t lb = tbase + ((uint32_t)(warp * 32) << 16);
        for (int k0 = 0; k0 < kp; k0 += 4) {
            const float4 w = *reinterpret_cast<const float4*>(wr + k0);
            const float h0 = b3_tf32_hi(w.x), h1 = b3_tf32_hi(w.y), h2 = b3_tf32_hi(w.z), h3 = b3_tf32_hi(w.w);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lb + k0), "r"(__float_as_uint(h0)),
                         "r"(__float_as_uint(h1)), "r"(__float_as_uint(h2)), "r"(__float_as_uint(h3)) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lb + kp + k0), "r"(__float_as_uint(w.x - h0)),
                         "r"(__float_as_uint(w.y - h1)), "r"(__float_as_uint(w.z - h2)), "r"(__float_as_uint(w.w - h3)) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (umma) {
        b3_tc_fence_before();
        __syncthreads();
        b3_tc_fence_after();
    }

    const int nq = P.n_it * M;
    const uint32_t tile_bytes = umma ? (uint32_t)(2 * utile * sizeof(float)) : (uint32_t)(D.kdim * NC * sizeof(float));
    const int xbuf_f = umma ? 2 * utile : kp * NC;          // floats of one input buffer
    const float* tile_in = P.tiles + D.in_off;
    float* tile_out = P.tiles + D.out_off;
    const unsigned int* cnt_in = P.cnt + (size_t)D.cnt_in * M * B3_CNT_STRIDE;
    unsigned int* cnt_out = P.cnt + (size_t)D.cnt_out * M * B3_CNT_STRIDE;

    if (ubuild && warp < 4) {
        // =============================== cell 0, tensor-core mode: the drain warps also evaluate the causal input stage
        // softsign(causal(wav_conv(samples))) (cswnv_shift1.py:352) straight into the B-operand buffer, from the samples
        // the tail published.  Two independent progress counters (tiles built / tiles drained) and non-blocking checks:
        // waiting for the next slot's samples must never hold back the drain of the current accumulator.
        constexpr int NA = SEG + 7;
        const uint32_t lb = tbase + ((uint32_t)(warp * 32) << 16) + 2 * kp;
        const float* itab = smem + D.sm_build;
        const float* cst = itab + K * 2 * H;
        const float* ibias = cst + H;
        float* av = smem + D.sm_build + K * 2 * H + 2 * H + warp * (BS * NA);      // per-warp copy of the slot's taps
        auto next_active = [&](int raw) {
            for (; raw < nq; ++raw) {
                const int it2 = raw / M;
                if (it2 - P.it0 < __ldg(P.slot_nit + (raw - it2 * M))) break;
            }
            return raw;
        };
        int qb = 0, qd = 0;                               // tiles built / drained so far
        int rb = next_active(0), rd = rb;                 // their positions in the ring order
        uint32_t idle = 0;
        while (rd < nq) {
            bool progress = false;
            if (qd < qb && b3_mbar_try(&d_full[DB(qd)], DU(qd) & 1)) {
                const int db = DB(qd);
                b3_tc_fence_after();
                if (qd > 0 && !b3_wait_warp(o_empty, (qd - 1) & 1, P.status)) return;
                for (int c0 = 0; c0 < NC; c0 += 16) {
                    uint32_t v[16];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(lb + db * NP + c0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (tid < RP) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < NC) Os[tid * OS + c0 + i] = __uint_as_float(v[i]);
                    }
                }
                b3_tc_fence_before();
                __syncwarp();
                if (lane == 0) { b3_mbar_arrive(&d_empty[db]); b3_mbar_arrive(o_full); }
                ++qd;
                rd = next_active(rd + 1);
                progress = true;
            }
            if (rb < nq && qb < qd + nxb) {
                const int it = rb / M, m = rb - it * M, buf = XB(qb);
                int ready = 1;
                if (lane == 0) {
                    if (qb >= nxb && !b3_mbar_try(&x_empty[buf], (XU(qb) - 1) & 1)) ready = 0;
                    if (ready) {
                        unsigned int v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt_in + (size_t)m * B3_CNT_STRIDE) : "memory");
                        ready = v >= (unsigned)(it + 1);
                    }
                }
                ready = __shfl_sync(0xffffffffu, ready, 0);
                if (ready) {
                    if (qb >= nxb && !b3_wait_warp(&x_empty[buf], (XU(qb) - 1) & 1, P.status)) return;     // every lane observes it
                    if (tid == 0) B3_PROF(0);
                    const int t0 = P.t_first + it * SEG;
                    for (int e = lane; e < BS * NA; e += 32) {
                        const int b = e / NA, i = e - b * NA;
                        const int n = t0 - (K - 1) + i - P.R - 1;           // newest sample seen by hidden time t0-(K-1)+i
                        av[e] = (i < SEG + K - 1 && n >= -SEG) ? __ldcg(P.hist_g + ((size_t)m * BS + b) * B3_HIST + ((n + 64 * B3_HIST) % B3_HIST)) : 0.f;
                    }
                    __syncwarp();
                    if (tid == 0) B3_PROF(7);
                    float* X = Xs + buf * xbuf_f;
                    const bool allv = t0 - (K - 1) >= 0;
                    for (int e = tid; e < (H >> 2) * NC; e += 128) {
                        const int kq = e / NC, n = e - kq * NC;
                        const int s2 = n / BS, b = n - s2 * BS;
                        const int r0 = kq * 4;
                        float acc[4];
                        if (allv) {
                            const float4 c0 = *reinterpret_cast<const float4*>(cst + r0);
                            acc[0] = c0.x; acc[1] = c0.y; acc[2] = c0.z; acc[3] = c0.w;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (j < K) {
                                    const float a = av[b * NA + s2 + j];
                                    const float4 cv = *reinterpret_cast<const float4*>(itab + (j * 2 + 1) * H + r0);
                                    acc[0] = fmaf(a, cv.x, acc[0]); acc[1] = fmaf(a, cv.y, acc[1]);
                                    acc[2] = fmaf(a, cv.z, acc[2]); acc[3] = fmaf(a, cv.w, acc[3]);
                                }
                        } else {
                            const float4 bi = *reinterpret_cast<const float4*>(ibias + r0);
                            acc[0] = bi.x; acc[1] = bi.y; acc[2] = bi.z; acc[3] = bi.w;
                            for (int j = 0; j < K; ++j)
                                if (t0 + s2 - (K - 1) + j >= 0) {
                                    const float a = av[b * NA + s2 + j];
                                    const float4 cb = *reinterpret_cast<const float4*>(itab + (j * 2) * H + r0);
                                    const float4 cv = *reinterpret_cast<const float4*>(itab + (j * 2 + 1) * H + r0);
                                    acc[0] += cb.x + cv.x * a; acc[1] += cb.y + cv.y * a;
                                    acc[2] += cb.z + cv.z * a; acc[3] += cb.w + cv.w * a;
                                }
                        }
                        const bool tv = t0 + s2 >= 0;
                        const float v0 = tv ? softsignf_(acc[0]) : 0.f, v1 = tv ? softsignf_(acc[1]) : 0.f;
                        const float v2 = tv ? softsignf_(acc[2]) : 0.f, v3 = tv ? softsignf_(acc[3]) : 0.f;
                        const float4 vh = make_float4(b3_tf32_hi(v0), b3_tf32_hi(v1), b3_tf32_hi(v2), b3_tf32_hi(v3));
                        const int xo = b3_uoff(n, r0, NP);
                        *reinterpret_cast<float4*>(X + xo) = vh;
                        *reinterpret_cast<float4*>(X + utile + xo) = make_float4(v0 - vh.x, v1 - vh.y, v2 - vh.z, v3 - vh.w);
                    }
                    if (tid == 0) B3_PROF(8);
                    b3_fence_proxy_async();          // generic writes -> the tensor core's (async proxy) reads
                    __syncwarp();
                    if (lane == 0) { b3_mbar_arrive(&x_full[buf]); b3_mbar_arrive(&x_lo[buf]); }
                    if (tid == 0) B3_PROF(9);
                    ++qb;
                    rb = next_active(rb + 1);
                    progress = true;
                }
            }
            if (!progress && ((++idle) & 255u) == 0) {
                int dead = ((volatile const int32_t*)P.status)[0] != 0;
                if (__any_sync(0xffffffffu, dead)) return;
            }
        }
        return;
    }
    if (umma && warp < 4) {
        // =============================== drain warps: accumulator tile (tensor memory) -> Os; thread = row
        const uint32_t lb = tbase + ((uint32_t)(warp * 32) << 16) + 2 * kp;
        B3_FOR_PASSES(q, it, m) {
            const int db = DB(q);
            if (!b3_wait_warp(&d_full[db], DU(q) & 1, P.status)) return;
            b3_tc_fence_after();
            if (tid == 0) B3_PROF(2);
            if (q > 0 && !b3_wait_warp(o_empty, (q - 1) & 1, P.status)) return;
            for (int c0 = 0; c0 < NC; c0 += 16) {
                uint32_t v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(lb + db * NP + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (tid < RP) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < NC) Os[tid * OS + c0 + i] = __uint_as_float(v[i]);
                }
            }
            b3_tc_fence_before();
            __syncwarp();
            if (lane == 0) { b3_mbar_arrive(&d_empty[db]); b3_mbar_arrive(o_full); }
        }
        return;
    }
    if (umma && warp == B3_EPI0) {
        // =============================== MMA issuer: D = W_hi.X_hi + W_lo.X_hi + W_hi.X_lo (fp32-equivalent, 3 x tf32)
        if (lane != 0) return;
        // instruction descriptor: D fp32, A/B tf32, both K-major, N = NP, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int kstep_bytes = 2 * (NP >> 3) * 128;                 // 8 channels = two core matrices along k
        B3_FOR_PASSES(q, it, m) {
            const int buf = XB(q), db = DB(q);
            if (!b3_mbar_wait(&x_full[buf], XU(q) & 1, P.status)) return;
            B3_PROF(1);
            if (q >= ndb && !b3_mbar_wait(&d_empty[db], (DU(q) - 1) & 1, P.status)) return;
            b3_tc_fence_after();
            const uint32_t xs = b3_smem(Xs + buf * xbuf_f);
            const uint32_t d_tmem = tbase + 2 * kp + db * NP;
            // the two passes that only read the high half of the activations start as soon as it has landed.
            // (One thread issues all 72 MMAs: the descriptor of a k-step is the previous one plus a constant in its address
            // field, and the loops are unrolled - built from scratch per MMA, the integer work between two issues was as long
            // as the instruction's own issue period of ~49 cycles, tools/microbench/umma_rate.cu.)
            const uint64_t dstep = (uint64_t)(kstep_bytes >> 4);
            uint64_t dh = b3_udesc(xs, NP);
            uint32_t a_hi = tbase, a_lo = tbase + kp;
            b3_umma_tf32_ts(d_tmem, a_hi, dh, idesc, 0);
            b3_umma_tf32_ts(d_tmem, a_lo, dh, idesc, 1);
#pragma unroll 4
            for (int ks = 1; ks < (kp >> 3); ++ks) {
                dh += dstep; a_hi += 8; a_lo += 8;
                b3_umma_tf32_acc(d_tmem, a_hi, dh, idesc);
                b3_umma_tf32_acc(d_tmem, a_lo, dh, idesc);
            }
            if (!b3_mbar_wait(&x_lo[buf], XU(q) & 1, P.status)) return;
            b3_tc_fence_after();
            uint64_t dl = b3_udesc(xs + utile * 4, NP);
            a_hi = tbase;
#pragma unroll 4
            for (int ks = 0; ks < (kp >> 3); ++ks) {
                b3_umma_tf32_acc(d_tmem, a_hi, dl, idesc);
                dl += dstep; a_hi += 8;
            }
            b3_umma_commit(&d_full[db]);
            b3_umma_commit(&x_empty[buf]);
        }
        return;
    }
    if (warp < B3_NGW) {
        // =============================== GEMM warps: warp = (row-block slot, k-split)
        const int slot = warp % SLOTS, ks = warp / SLOTS;
        if (slot >= nslot) return;
        const int rg = lane / NCG, cg = lane - rg * NCG;
        const bool lane_ok = rg < RGW;
        const int kchunk = (kp + KS - 1) / KS;
        const int kb = ks * kchunk, ke = min(kp, kb + kchunk);
        uint32_t kphase = 0;
        B3_FOR_PASSES(q, it, m) {
            const int buf = XB(q);
            if (!b3_wait_warp(&x_full[buf], XU(q) & 1, P.status)) return;
            if (tid == 0) B3_PROF(1);
            const float* X = Xs + buf * kp * NC;
            bool first = true;
            for (int rb = slot; rb < nrb; rb += nslot) {
                const int r0 = (rb * RGW + rg) * TR;
                const bool act = lane_ok && r0 < RP;
                float acc[TR][TC];
#pragma unroll
                for (int i = 0; i < TR; ++i)
#pragma unroll
                    for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;
                if (act) {
                    const float* wp = Wt + r0;
                    const float* xp = X + cg * TC;
#pragma unroll 2
                    for (int k = kb; k < ke; ++k) {
                        float x[TC];
                        if constexpr (TC % 4 == 0) {
#pragma unroll
                            for (int j = 0; j < TC; j += 4) {
                                const float4 v = *reinterpret_cast<const float4*>(xp + k * NC + j);
                                x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
                            }
                        } else if constexpr (TC % 2 == 0) {
#pragma unroll
                            for (int j = 0; j < TC; j += 2) {
                                const float2 v = *reinterpret_cast<const float2*>(xp + k * NC + j);
                                x[j] = v.x; x[j + 1] = v.y;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < TC; ++j) x[j] = xp[k * NC + j];
                        }
                        float w[TR];
                        if constexpr (TR % 4 == 0) {
#pragma unroll
                            for (int i = 0; i < TR; i += 4) {
                                const float4 v = *reinterpret_cast<const float4*>(wp + (size_t)k * RP + i);
                                w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < TR; i += 2) {
                                const float2 v = *reinterpret_cast<const float2*>(wp + (size_t)k * RP + i);
                                w[i] = v.x; w[i + 1] = v.y;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < TR; ++i)
#pragma unroll
                            for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
                    }
                }
                if (first) {
                    first = false;
                    if (tid == 0) B3_PROF(2);
                    if (q > 0 && !b3_wait_warp(o_empty, (q - 1) & 1, P.status)) return;
                }
                if (KS == 1 || ks == 0) {
                    if (KS > 1 && rb != slot && !b3_wait_warp(&kback[slot], (kphase - 1) & 1, P.status)) return;
                    if (act) {
#pragma unroll
                        for (int i = 0; i < TR; ++i)
#pragma unroll
                            for (int j = 0; j < TC; ++j) Os[(r0 + i) * OS + cg * TC + j] = acc[i][j];
                    }
                    if (KS > 1) {
                        __syncwarp();
                        if (lane == 0) b3_mbar_arrive(&kbar[slot]);
                    }
                } else {
                    // second k-split: add to what the first one stored (fixed order: deterministic)
                    if (!b3_wait_warp(&kbar[slot], kphase & 1, P.status)) return;
                    if (act) {
#pragma unroll
                        for (int i = 0; i < TR; ++i)
#pragma unroll
                            for (int j = 0; j < TC; ++j) Os[(r0 + i) * OS + cg * TC + j] += acc[i][j];
                    }
                    __syncwarp();
                    if (lane == 0) b3_mbar_arrive(&kback[slot]);
                }
                ++kphase;
            }
            __syncwarp();
            if (lane == 0) { b3_mbar_arrive(o_full); b3_mbar_arrive(&x_empty[buf]); }
        }
        return;
    }

    if (MODE == 0 && build && warp >= B3_EPI0 && warp < p1w0) {
        // =============================== builders (cell 0): softsign(causal(wav_conv(samples))) straight into the
        // GEMM operand buffer (cswnv_shift1.py:352); a thread owns channel r of utterance b, whose SEG columns share taps
        constexpr int NA = SEG + 7;
        const int tb = tid - B3_EPI0 * 32;
        const float* itab = smem + D.sm_build;
        const float* cst = itab + K * 2 * H;
        const float* ibias = cst + H;
        float* av = smem + D.sm_build + K * 2 * H + 2 * H;          // [2][BS][NA]
        uint32_t bp = 0;
        B3_FOR_PASSES(q, it, m) {
            const int buf = XB(q);
            const int t0 = P.t_first + it * SEG;
            if (q >= nxb && !b3_wait_warp(&x_empty[buf], (XU(q) - 1) & 1, P.status)) return;
            int ok = 1;
            if (lane == 0) ok = b3_cnt_wait(cnt_in + (size_t)m * B3_CNT_STRIDE, (unsigned)(it + 1), P, 100) ? 1 : 0;
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (!ok) return;
            if (tb == 0) B3_PROF(0);
            float* a = av + buf * BS * NA;
            for (int e = tb; e < BS * NA; e += B3_NBW * 32) {
                const int b = e / NA, i = e - b * NA;
                const int n = t0 - (K - 1) + i - P.R - 1;               // newest sample seen by hidden time t0-(K-1)+i
                a[e] = (i < SEG + K - 1 && n >= -SEG) ? __ldcg(P.hist_g + ((size_t)m * BS + b) * B3_HIST + ((n + 64 * B3_HIST) % B3_HIST)) : 0.f;
            }
            __syncwarp();
            if (lane == 0) b3_mbar_arrive(bbar);
            if (!b3_wait_warp(bbar, bp & 1, P.status)) return;
            ++bp;
            float* X = Xs + buf * kp * NC;
            const bool allv = t0 - (K - 1) >= 0;
            for (int e = tb; e < H * BS; e += B3_NBW * 32) {
                const int r = e / BS, b = e - r * BS;
                float acc[SEG];
                if (allv) {
                    const float c = cst[r];
#pragma unroll
                    for (int s2 = 0; s2 < SEG; ++s2) acc[s2] = c;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < K) {
                            const float cv = itab[(j * 2 + 1) * H + r];
#pragma unroll
                            for (int s2 = 0; s2 < SEG; ++s2) acc[s2] = fmaf(a[b * NA + s2 + j], cv, acc[s2]);
                        }
                    }
#pragma unroll
                    for (int s2 = 0; s2 < SEG; ++s2) X[r * NC + s2 * BS + b] = softsignf_(acc[s2]);
                } else {
                    for (int s2 = 0; s2 < SEG; ++s2) {
                        float v = 0.f;
                        if (t0 + s2 >= 0) {
                            v = ibias[r];
                            for (int j = 0; j < K; ++j)
                                if (t0 + s2 - (K - 1) + j >= 0) v += itab[(j * 2) * H + r] + itab[(j * 2 + 1) * H + r] * a[b * NA + s2 + j];
                            v = softsignf_(v);
                        }
                        X[r * NC + s2 * BS + b] = v;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) b3_mbar_arrive(&x_full[buf]);
        }
        return;
    }

    if (warp == B3_LOADER) {
        // =============================== loader
        if (lane != 0 || build || ubuild) return;
        B3_FOR_PASSES(q, it, m) {
            const int buf = XB(q);
            if (q >= nxb && !b3_mbar_wait(&x_empty[buf], (XU(q) - 1) & 1, P.status)) return;
            if (!b3_cnt_wait(cnt_in + (size_t)m * B3_CNT_STRIDE, (unsigned)D.nprod_in * (unsigned)(it + 1), P, 100 + MODE * 50 + l)) return;
            B3_PROF(0);
            b3_fence_proxy_async();
            if (umma) {
                // each half as B3_TMA_SPLIT bulk copies (measured: 8 concurrent pieces are not faster than one copy, the SM pulls ~24 B/clk from L2 either way)
                const uint32_t half = tile_bytes >> 1;
                const uint32_t piece = ((half / B3_TMA_SPLIT) + 15u) & ~15u;
                b3_mbar_arm(&x_full[buf], half);
                for (uint32_t o = 0; o < half; o += piece)
                    b3_tma_load(Xs + buf * xbuf_f + (o >> 2), tile_in + (size_t)m * D.in_slot_f + (o >> 2), min(piece, half - o), &x_full[buf]);
                b3_mbar_arm(&x_lo[buf], half);
                for (uint32_t o = 0; o < half; o += piece)
                    b3_tma_load(Xs + buf * xbuf_f + utile + (o >> 2), tile_in + (size_t)m * D.in_slot_f + utile + (o >> 2), min(piece, half - o), &x_lo[buf]);
            } else {
                b3_mbar_arm(&x_full[buf], tile_bytes);
                b3_tma_load(Xs + buf * xbuf_f, tile_in + (size_t)m * D.in_slot_f, tile_bytes, &x_full[buf]);
            }
        }
        return;
    }

    // =================================== epilogue warps
    const int te = tid - epi0 * 32;
    const float* bias = P.wbuf + D.b_off;
    uint32_t ep = 0;
    auto epi_sync = [&]() -> bool {
        __syncwarp();
        if (lane == 0) b3_mbar_arrive(ebar);
        const bool ok = b3_wait_warp(ebar, ep & 1, P.status);
        ++ep;
        return ok;
    };

    if constexpr (MODE == 1) {
        // ---- dense head stage: out[row][b] = act(W.x + bias (+ the skips of layers 0..L-2))
        const int nout = D.nrows * BS;
        constexpr bool V4 = BS % 4 == 0;
        constexpr int VW = V4 ? 4 : 1;                  // outputs per item
        const int nitem = nout / VW;
        B3_FOR_PASSES(q, it, m) {
            const int buf = XB(q);
            if (!b3_wait_warp(&x_full[buf], XU(q) & 1, P.status)) return;      // the skips are complete once the input is
            // running sum of the skips of layers 0..L-2 (chained through the cells): up to two items per thread stay
            // in registers while the GEMM runs, the rest (small CTAs counts / odd shapes) goes through shared memory
            const bool sk = D.add_skips && P.L >= 2;
            const float* skp = P.skipbuf + ((size_t)(P.L - 2) * M + m) * P.S * BS + (size_t)D.row0 * BS;
            float pr[4][VW];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = te + u * nepi;
#pragma unroll
                for (int v = 0; v < VW; ++v) pr[u][v] = 0.f;
                if (sk && e < nitem) {
                    if constexpr (V4) { const float4 x = __ldcg(reinterpret_cast<const float4*>(skp) + e); pr[u][0] = x.x; pr[u][1] = x.y; pr[u][2] = x.z; pr[u][3] = x.w; }
                    else pr[u][0] = __ldcg(skp + e);
                }
            }
            for (int e = te + 4 * nepi; e < nitem; e += nepi)
#pragma unroll
                for (int v = 0; v < VW; ++v) pre1[e * VW + v] = sk ? __ldcg(skp + e * VW + v) : 0.f;
            if (!b3_wait_warp(o_full, q & 1, P.status)) return;
            if (te == 0) B3_PROF(4);
            float* tout = tile_out + (size_t)m * D.out_slot_f + (size_t)D.row0 * BS;
            for (int e = te, u = 0; e < nitem; e += nepi, ++u) {
                const int r = (e * VW) / BS, b = e * VW - r * BS;
                float o[VW];
#pragma unroll
                for (int v = 0; v < VW; ++v) {
                    float x = bias[r] + Os[r * OS + b + v] + (u == 0 ? pr[0][v] : u == 1 ? pr[1][v] : u == 2 ? pr[2][v] : u == 3 ? pr[3][v] : pre1[e * VW + v]);
                    o[v] = D.relu ? fmaxf(x, 0.f) : x;
                }
                if constexpr (V4) __stcg(reinterpret_cast<float4*>(tout) + e, make_float4(o[0], o[1], o[2], o[3]));
                else __stcg(tout + e, o[0]);
            }
            if (!epi_sync()) return;
            if (te == 0) {
                b3_signal(cnt_out + (size_t)m * B3_CNT_STRIDE);
                b3_mbar_arrive(o_empty); b3_mbar_arrive(&x_empty[buf]);
                B3_PROF(5);
            }
        }
        return;
    } else {
    float* ringS = smem + D.sm_ring;
    float* ringG = P.ring_scratch + D.ring_off;
    const int twoH = 2 * H;
    const int nitems = np * NC;
    // ---- P1: everything the gate needs that does not depend on the GEMM (conditioning, ring of partial sums, bias)
    auto p1 = [&](const int it, const int m, float4* dst, const int t_id, const int t_n) {
        const int t0 = P.t_first + it * SEG;
        for (int item = t_id; item < nitems; item += t_n) {
            const int pi = item / NC, col = item - pi * NC;
            const int s = col / BS, b = col - s * BS;
            const int t = t0 + s;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0) {
                int bglob = m * BS + b; if (bglob >= P.B) bglob = P.B - 1;
                const int i = D.pair0 + pi;
                float az = __ldg(P.pconst + l * twoH + i), ac = __ldg(P.pconst + l * twoH + H + i);
                // frame / phase of tau = t + sp (replicate pad before R): one division, then incremental
                int d = t - P.R, f = 0, j = 0;
                if (d > 0) { f = d / P.U; j = d - f * P.U; }
                const float* pb = P.P + ((size_t)bglob * P.Tf * P.L + l) * SEG * twoH + i;
                const size_t fstride = (size_t)P.L * SEG * twoH;
#pragma unroll
                for (int sp = 0; sp < SEG; ++sp) {
                    const float g = __ldg(P.gain + j);
                    const float* pr = pb + (size_t)f * fstride + sp * twoH;
                    az = fmaf(g, __ldg(pr), az);
                    ac = fmaf(g, __ldg(pr + H), ac);
                    ++d;
                    if (d > 0 && ++j == P.U) { j = 0; ++f; }
                }
                const int slot = (t - P.t_first) % RL;
                const size_t ro = ((size_t)slot * np2 + pi * 2) * Bp + m * BS + b;
                float rz, rc;
                if (D.ring_in_smem) { rz = ringS[ro]; rc = ringS[ro + Bp]; ringS[ro] = 0.f; ringS[ro + Bp] = 0.f; }
                else { rz = __ldcg(ringG + ro); rc = __ldcg(ringG + ro + Bp); __stcg(ringG + ro, 0.f); __stcg(ringG + ro + Bp, 0.f); }
                v = make_float4(az, ac, rz + bias[pi * 2], rc + bias[pi * 2 + 1]);
            }
            dst[item] = v;
        }
    };
    if (psplit && warp < epi0) {
        // =============================== prefetch team: one pass ahead of the gate.  pre_empty is released after P3 of the
        // pass two behind, so the ring slots read (and cleared) here are never ones a running P3 still adds to (M >= 2)
        const int tp = tid - p1w0 * 32;
        int m_prev = -1;
        B3_FOR_PASSES(q, it, m) {
            const int pb = q & 1;
            if (q >= 2 && !b3_wait_warp(&pre_empty[pb], ((q >> 1) - 1) & 1, P.status)) return;
            // a ragged batch can leave a single slot in the ring: then the previous pass is the same utterances and its
            // scatter into the ring must be complete before the ring is read here
            if (q >= 1 && m == m_prev && !b3_wait_warp(&pre_empty[pb ^ 1], ((q - 1) >> 1) & 1, P.status)) return;
            m_prev = m;
            p1(it, m, pre + pb * nitems, tp, B3_P1W * 32);
            __syncwarp();
            if (lane == 0) b3_mbar_arrive(&pre_full[pb]);
        }
        return;
    }
    B3_FOR_PASSES(q, it, m) {
        const int buf = XB(q);
        const int t0 = P.t_first + it * SEG;
        float4* prq = pre + (psplit ? (q & 1) * nitems : 0);
        if (!psplit) p1(it, m, prq, te, nepi);
        if (te == 0) B3_PROF(3);
        if (!b3_wait_warp(&x_full[buf], XU(q) & 1, P.status)) return;
        if (umma && !b3_wait_warp(&x_lo[buf], XU(q) & 1, P.status)) return;
        // running sum of the skips of the layers below (complete once the input is): fetch it while the GEMM runs
        const bool sk_pre = D.nskip * BS <= nepi;
        float sk_prev = 0.f;
        if (l >= 2 && sk_pre && te < D.nskip * BS)
            sk_prev = __ldcg(P.skipbuf + ((size_t)(l - 2) * M + m) * P.S * BS + (size_t)D.skip0 * BS + te);
        if (!b3_wait_warp(o_full, q & 1, P.status)) return;
        if (psplit && !b3_wait_warp(&pre_full[q & 1], (q >> 1) & 1, P.status)) return;
        if (te == 0) B3_PROF(4);
        const float* X = Xs + buf * xbuf_f;
        float* tout = tile_out + (size_t)m * D.out_slot_f;
        // ---- P2: gate (cswnv_shift1.py:281-285) and publish
        for (int item = te; item < nitems; item += nepi) {
            const int pi = item / NC, col = item - pi * NC;
            const int s = col / BS, b = col - s * BS;
            const int t = t0 + s;
            float hv = 0.f;
            if (t >= 0) {
                const float4 v = prq[item];
                float qz = v.z, qc = v.w;
                for (int j = K - 1; j >= 0; --j) {
                    const int delta = (K - 1 - j) * dil;
                    if (delta > s) break;
                    const float* o = Os + ((j * np + pi) * 2) * OS + (s - delta) * BS + b;
                    qz += o[0]; qc += o[OS];
                }
                // sigmoid / tanh through ex2.approx + rcp (abs. error ~1e-7, inside the 1e-4 head tolerance)
                const float gz = qz * v.x, gc = qc * v.y;
                const float ez = __expf(-gz), ec = __expf(-2.f * fabsf(gc));
                const float z = __frcp_rn(1.f + ez);
                const float th = copysignf((1.f - ec) * __frcp_rn(1.f + ec), gc);
                float carry;
                if (umma) { const int xo = b3_uoff(col, D.pair0 + pi, NP); carry = X[xo] + X[utile + xo]; }
                else carry = X[(D.pair0 + pi) * NC + col];
                hv = (1.f - z) * th + z * carry;
            }
            if (B3_ABL == 2) {
            } else if (!D.is_last) {
                if (umma) {      // exact split into tf32 halves for the next cell's tensor-core passes
                    const int xo = b3_uoff(col, D.pair0 + pi, NP);
                    const float hh = b3_tf32_hi(hv);
                    __stcg(tout + xo, hh); __stcg(tout + utile + xo, hv - hh);
                } else __stcg(tout + (size_t)(D.pair0 + pi) * NC + col, hv);
            } else if (s == SEG - 1) __stcg(tout + (size_t)(D.pair0 + pi) * BS + b, hv);     // the head only needs the last column
        }
        // rows of out_skip[l-1] applied to the last time column of every utterance, added to the running sum
        if (l >= 1) {
            float* sk = P.skipbuf + ((size_t)(l - 1) * M + m) * P.S * BS + (size_t)D.skip0 * BS;
            for (int e = te; e < D.nskip * BS; e += nepi) {
                const int r = e / BS, b = e - r * BS;
                float v = Os[(K * np2 + r) * OS + (SEG - 1) * BS + b];
                if (l >= 2) v += sk_pre ? sk_prev : __ldcg(P.skipbuf + ((size_t)(l - 2) * M + m) * P.S * BS + (size_t)D.skip0 * BS + e);
                __stcg(sk + e, v);
            }
        }
        if (!epi_sync()) return;
        if (te == 0) {
            b3_signal(cnt_out + (size_t)m * B3_CNT_STRIDE);
            b3_mbar_arrive(&x_empty[buf]);          // the carry has been read: P3 below only needs the GEMM output
            B3_PROF(5);
        }
        // ---- P3 (off the critical path): delayed taps into the ring of partial sums; the BS utterances of one
        // (tap, row, time column) are contiguous in the GEMM output and in the ring
        {
            const bool uniq = dil >= SEG;               // every (tap, column) has its own target slot
            const int nt = uniq ? (K - 1) * SEG : (K - 1) * dil;
            const int n3 = nt * np2;
            for (int e = te; e < n3; e += nepi) {
                const int r2 = e % np2, tt = e / np2;
                float val[BS];
                int trel;
                if (uniq) {
                    const int s = tt % SEG, j = tt / SEG;                    // tap 0..K-2
                    trel = s + (K - 1 - j) * dil;
                    const float* o = Os + (j * np2 + r2) * OS + s * BS;
#pragma unroll
                    for (int b = 0; b < BS; ++b) val[b] = o[b];
                } else {
                    // several (tap, column) pairs share a target: summed by one thread in a fixed order
                    trel = SEG + tt;
#pragma unroll
                    for (int b = 0; b < BS; ++b) val[b] = 0.f;
                    for (int j = K - 2; j >= 0; --j) {
                        const int s = trel - (K - 1 - j) * dil;
                        if (s < 0) break;
                        if (s < SEG) {
                            const float* o = Os + (j * np2 + r2) * OS + s * BS;
#pragma unroll
                            for (int b = 0; b < BS; ++b) val[b] += o[b];
                        }
                    }
                }
                const int slot = (t0 + trel - P.t_first) % RL;
                const size_t ro = ((size_t)slot * np2 + r2) * Bp + m * BS;
                if (D.ring_in_smem) {
#pragma unroll
                    for (int b = 0; b < BS; ++b) ringS[ro + b] += val[b];
                } else if (B3_ABL == 1) {
                } else if constexpr (BS % 4 == 0) {
#pragma unroll
                    for (int b = 0; b < BS; b += 4)                         // RED: fire and forget, one writer per element
                        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ringG + ro + b),
                                     "f"(val[b]), "f"(val[b + 1]), "f"(val[b + 2]), "f"(val[b + 3]) : "memory");
                } else {
#pragma unroll
                    for (int b = 0; b < BS; ++b) atomicAdd(ringG + ro + b, val[b]);
                }
            }
        }
        if (!epi_sync()) return;
        if (te == 0) { b3_mbar_arrive(o_empty); if (psplit) b3_mbar_arrive(&pre_empty[q & 1]); B3_PROF(6); }
    }
    }
}

// ------------------------------------------------------------------ tail CTA
// Serves whole slots (m = hindex, hindex + NT, ...): [out_2 when it is small] -> sampler -> causal input stage
// of the next step.  The noise of the step is drawn before the wait for the head (off the critical path).
// NTT threads form one tail instance; a tail CTA runs one instance of 512 threads or two of 256 side by side (the
// phases of a tail are short dependent chains: two instances on different slots hide each other's latency).
template <int SEG, int BS, int NTT>
__device__ void b3_tail(const BatchParams& P, const BCta& D, float* smem_cta, const int inst, const int ninst) {
    constexpr int NC = SEG * BS;
    const int tid = threadIdx.x - inst * NTT, lane = tid & 31, warp = tid >> 5;
    float* smem = smem_cta + (size_t)inst * D.tail_stride;
    const int NVT = P.NT * ninst;                    // tail instances of the whole grid
    auto tsync = [inst]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + inst), "n"(NTT) : "memory"); };
    const int H = P.H, K = P.K, n1 = P.n1, n1p = P.n1p, n2 = P.n2, n2p = P.n2p, M = P.M;
    const int hc = D.hindex * ninst + inst;
    const bool is_c = P.kind == 0;

    float* x2 = smem + D.sm_x2;                     // tail_gemm: [BS][n1p] relu(out_1), utterance-major
    float* oc = smem + D.sm_oc;                     // [BS][n2p] out_2 rows
    float* w2 = smem + D.sm_w2;                     // tail_gemm: [n2][n1p]
    float* tin = smem + D.sm_tin;                   // [H][NC]
    float* itab = smem + D.sm_itab;                 // CSWNV: in_tab [K][2][H]; then (both kinds) in_bias [H], out_2 bias [n2]
    float* ibias = itab + (is_c ? K * 2 * H : 0);
    float* cst = ibias + H;                         // CSWNV: in_bias[r] + sum_j cb_j[r]
    float* b2s = cst + H;
    float2* ain = reinterpret_cast<float2*>(smem + D.sm_ain);     // [NC][K]
    float* nz = smem + D.sm_noise;                  // CSWNV: eps [BS][SEG]; DSWNV: -log q [BS][Q]
    float* hist = smem + D.sm_hist;                 // [MO][BS][HIST]
    float* lpb = smem + D.sm_lpb;                   // [MO][BS][8]
    int* flag = reinterpret_cast<int*>(smem + D.sm_misc);

    const int nown = hc < M ? (M - hc + NVT - 1) / NVT : 0;       // slots served: m = hc + i*NT

    for (int i = tid; i < BS * (n1p + 4); i += NTT) x2[i] = 0.f;
    for (int i = tid; i < nown * BS * B3_HIST; i += NTT) hist[i] = is_c ? 0.f : __int_as_float(P.Q / 2);
    for (int i = tid; i < nown * BS * 8; i += NTT) lpb[i] = 0.f;
    if (is_c)
        for (int i = tid; i < K * 2 * H; i += NTT) itab[i] = P.in_tab[i];
    for (int i = tid; i < H; i += NTT) ibias[i] = P.in_bias[i];
    for (int i = tid; i < n2; i += NTT) b2s[i] = P.wbuf[P.b2_off + i];
    if (is_c)
        for (int r = tid; r < H; r += NTT) {
            float c = P.in_bias[r];
            for (int j = 0; j < K; ++j) c += P.in_tab[(j * 2) * H + r];
            cst[r] = c;
        }
    if (P.tail_gemm)
        for (int i = tid; i < n2 * n1p; i += NTT) w2[i] = P.wbuf[P.w2_off + i];
    tsync();
    if (P.seedw) {
        for (int e = tid; e < nown * BS * SEG; e += NTT) {
            const int mb = e / SEG, j = e - mb * SEG;
            const int ml = mb / BS, b = mb - ml * BS;
            int bglob = (hc + ml * NVT) * BS + b; if (bglob >= P.B) bglob = P.B - 1;
            const float v = is_c ? ((const float*)P.seedw)[(size_t)bglob * SEG + j]
                                 : __int_as_float(pos_mod(((const long long*)P.seedw)[bglob], P.Q));
            hist[(size_t)mb * B3_HIST + ((j - SEG + B3_HIST) % B3_HIST)] = v;
            const int mm = P.lpc - SEG + j;
            if (is_c && mm >= 0) lpb[(size_t)mb * 8 + mm] = v;
        }
        tsync();
    }

    // CSWNV, tensor-core tile layout, steady state: thread (channel quad kq, column phase cg) keeps its seven tap
    // coefficients and constants in registers for the whole utterance and evaluates four columns per slot
    constexpr int NCQ = NC / 4;                      // column phases of one channel quad
    // column groups of 4 per thread and pass: as few as keep (H/4) x (NCQ/CPT) threads inside the instance at H = 192
    constexpr int CPT = NTT >= 512 ? (NCQ > 10 && NCQ % 2 == 0 ? 2 : 1)
                                   : (NCQ > 10 && NCQ % 4 == 0 ? 4 : (NCQ > 5 && NCQ % 2 == 0 ? 2 : 1));
    constexpr int TPQ = NCQ / CPT;                   // threads per channel quad
    const bool fastq = P.umma && is_c && K <= 8 && NC % 4 == 0 && (H >> 2) * TPQ <= NTT;
    float4 pcv[8], pcst = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) pcv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (fastq && tid < (H >> 2) * TPQ) {
        const int r0 = (tid / TPQ) * 4;
        pcst = *reinterpret_cast<const float4*>(cst + r0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < K) pcv[j] = *reinterpret_cast<const float4*>(itab + (j * 2 + 1) * H + r0);
    }
    // Input stage for iteration `itn` of slot m: h0 = softsign(causal(...)) (cswnv_shift1.py:352 / dswnv.py:345).
    // A thread owns one channel r of one utterance b: the SEG columns of that utterance share their taps.
    auto input_stage = [&](int itn, int m, int ml) {
        if (P.build_x) {                              // cell 0 builds its own input from the published samples
            tsync();
            const int it = itn - 1;
            if (tid == 0 && itn > 0) B3_PROF(6);
            if (tid == 0) b3_signal(P.cnt + (size_t)m * B3_CNT_STRIDE);
            return;
        }
        const int t0n = P.t_first + itn * SEG;
        const float* hs = hist + (size_t)ml * BS * B3_HIST;
        float* dst = P.tiles + (size_t)m * (P.umma ? 2 * H * P.NP : H * NC);      // layer 0 input (tile offset 0)
        if (fastq && t0n - (K - 1) >= 0) {
            constexpr int AST = (SEG + 7) | 1;                       // odd stride: conflict-free tap reads
            float* as = reinterpret_cast<float*>(ain);
            for (int e = tid; e < BS * (SEG + 7); e += NTT) {
                const int b = e / (SEG + 7), i = e - b * (SEG + 7);
                const int nn = t0n - (K - 1) + i - P.R - 1;           // newest sample seen by hidden time t0n-(K-1)+i
                as[b * AST + i] = (i < SEG + K - 1 && nn >= -SEG) ? hs[b * B3_HIST + ((nn + 64 * B3_HIST) % B3_HIST)] : 0.f;
            }
            tsync();
            if (tid < (H >> 2) * TPQ) {
                const int kq = tid / TPQ, cg = tid - kq * TPQ, r0 = kq * 4;
                const int utile = H * P.NP;
#pragma unroll
                for (int u = 0; u < 4 * CPT; ++u) {
                    const int n = u * TPQ + cg;                      // lanes write consecutive columns: coalesced 16-byte words
                    const int s2 = n / BS, b = n - s2 * BS;
                    float4 acc = pcst;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < K) {
                            const float a = as[b * AST + s2 + j];
                            acc.x = fmaf(a, pcv[j].x, acc.x); acc.y = fmaf(a, pcv[j].y, acc.y);
                            acc.z = fmaf(a, pcv[j].z, acc.z); acc.w = fmaf(a, pcv[j].w, acc.w);
                        }
                    // softsign with a correctly rounded reciprocal (<= 1 ulp from the division)
                    const float v0 = acc.x * __frcp_rn(1.f + fabsf(acc.x)), v1 = acc.y * __frcp_rn(1.f + fabsf(acc.y));
                    const float v2 = acc.z * __frcp_rn(1.f + fabsf(acc.z)), v3 = acc.w * __frcp_rn(1.f + fabsf(acc.w));
                    const float4 vh = make_float4(b3_tf32_hi(v0), b3_tf32_hi(v1), b3_tf32_hi(v2), b3_tf32_hi(v3));
                    const int xo = b3_uoff(n, r0, P.NP);
                    __stcg(reinterpret_cast<float4*>(dst + xo), vh);
                    __stcg(reinterpret_cast<float4*>(dst + utile + xo), make_float4(v0 - vh.x, v1 - vh.y, v2 - vh.z, v3 - vh.w));
                }
            }
        } else if (P.umma && K <= 8) {
            // tensor-core tile layout: a thread owns 4 consecutive channels of one column and writes them as one
            // 16-byte word (hi) + one (lo); 8 consecutive columns fill a 128-byte core matrix
            const int utile = H * P.NP;
            const bool allv = t0n - (K - 1) >= 0;
            for (int e = tid; e < (H >> 2) * NC; e += NTT) {
                const int kq = e / NC, n = e - kq * NC;
                const int s = n / BS, b = n - s * BS;
                const int r0 = kq * 4;
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int nn = t0n + s - (K - 1) + j - P.R - 1;         // newest sample seen by tap j of this column
                    a[j] = (j < K && nn >= -SEG) ? hs[b * B3_HIST + ((nn + 64 * B3_HIST) % B3_HIST)] : (is_c ? 0.f : __int_as_float(P.Q / 2));
                }
                float acc[4];
                if (is_c && allv) {
                    const float4 c0 = *reinterpret_cast<const float4*>(cst + r0);
                    acc[0] = c0.x; acc[1] = c0.y; acc[2] = c0.z; acc[3] = c0.w;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < K) {
                            const float4 cv = *reinterpret_cast<const float4*>(itab + (j * 2 + 1) * H + r0);
                            acc[0] = fmaf(a[j], cv.x, acc[0]); acc[1] = fmaf(a[j], cv.y, acc[1]);
                            acc[2] = fmaf(a[j], cv.z, acc[2]); acc[3] = fmaf(a[j], cv.w, acc[3]);
                        }
                } else {
                    const float4 bi = *reinterpret_cast<const float4*>(ibias + r0);
                    acc[0] = bi.x; acc[1] = bi.y; acc[2] = bi.z; acc[3] = bi.w;
                    if (is_c) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < K && t0n + s - (K - 1) + j >= 0) {
                                const float4 cb = *reinterpret_cast<const float4*>(itab + (j * 2) * H + r0);
                                const float4 cv = *reinterpret_cast<const float4*>(itab + (j * 2 + 1) * H + r0);
                                acc[0] += cb.x + cv.x * a[j]; acc[1] += cb.y + cv.y * a[j];
                                acc[2] += cb.z + cv.z * a[j]; acc[3] += cb.w + cv.w * a[j];
                            }
                    } else {
                        float4 g[8];                         // all gathers in flight before the first add
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            g[j] = (j < K && t0n + s - (K - 1) + j >= 0)
                                       ? __ldg(reinterpret_cast<const float4*>(P.in_tab + ((size_t)j * P.Q + __float_as_int(a[j])) * H + r0))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { acc[0] += g[j].x; acc[1] += g[j].y; acc[2] += g[j].z; acc[3] += g[j].w; }
                    }
                }
                float4 vh, vl;
                {
                    const bool tv = t0n + s >= 0;
                    const float v0 = tv ? softsignf_(acc[0]) : 0.f, v1 = tv ? softsignf_(acc[1]) : 0.f;
                    const float v2 = tv ? softsignf_(acc[2]) : 0.f, v3 = tv ? softsignf_(acc[3]) : 0.f;
                    vh = make_float4(b3_tf32_hi(v0), b3_tf32_hi(v1), b3_tf32_hi(v2), b3_tf32_hi(v3));
                    vl = make_float4(v0 - vh.x, v1 - vh.y, v2 - vh.z, v3 - vh.w);
                }
                const int xo = b3_uoff(n, r0, P.NP);
                __stcg(reinterpret_cast<float4*>(dst + xo), vh);
                __stcg(reinterpret_cast<float4*>(dst + utile + xo), vl);
            }
        } else if (K <= 8) {
            for (int e = tid; e < H * BS; e += NTT) {
                const int r = e / BS, b = e - r * BS;
                // samples / indices seen by hidden times t0n-(K-1) .. t0n+SEG-1 (newest sample of hidden time tau: tau-R-1)
                float a[SEG + 7];
#pragma unroll
                for (int i = 0; i < SEG + 7; ++i) {
                    const int n = t0n - (K - 1) + i - P.R - 1;
                    a[i] = (i < SEG + K - 1 && n >= -SEG) ? hs[b * B3_HIST + ((n + 64 * B3_HIST) % B3_HIST)]
                                                           : (is_c ? 0.f : __int_as_float(P.Q / 2));
                }
                const float bias = ibias[r];
                float acc[SEG];
#pragma unroll
                for (int s = 0; s < SEG; ++s) acc[s] = bias;
                if (is_c && t0n - (K - 1) >= 0) {            // every tap lies inside the sequence (all but the first steps)
                    const float c0 = cst[r];
#pragma unroll
                    for (int s = 0; s < SEG; ++s) acc[s] = c0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < K) {
                            const float cv = itab[(j * 2 + 1) * H + r];
#pragma unroll
                            for (int s = 0; s < SEG; ++s) acc[s] = fmaf(a[s + j], cv, acc[s]);
                        }
                    }
                } else if (is_c) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < K) {
                            const float cb = itab[(j * 2) * H + r], cv = itab[(j * 2 + 1) * H + r];
#pragma unroll
                            for (int s = 0; s < SEG; ++s) {
                                const float vld = (t0n + s - (K - 1) + j >= 0) ? 1.f : 0.f;      // conv zero padding
                                acc[s] = fmaf(vld, cb, fmaf(vld * a[s + j], cv, acc[s]));
                            }
                        }
                    }
                } else {
                    float g[SEG][8];                         // all gathers in flight before the first add
#pragma unroll
                    for (int s = 0; s < SEG; ++s)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            g[s][j] = (j < K && t0n + s - (K - 1) + j >= 0)
                                          ? __ldg(P.in_tab + ((size_t)j * P.Q + __float_as_int(a[s + j])) * H + r) : 0.f;
#pragma unroll
                    for (int s = 0; s < SEG; ++s)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[s] += g[s][j];
                }
#pragma unroll
                for (int s = 0; s < SEG; ++s) {
                    const float v = (t0n + s >= 0) ? softsignf_(acc[s]) : 0.f;
                    if (P.umma) {
                        const int xo = b3_uoff(s * BS + b, r, P.NP);
                        const float hh = b3_tf32_hi(v);
                        __stcg(dst + xo, hh); __stcg(dst + H * P.NP + xo, v - hh);
                    } else __stcg(dst + (size_t)r * NC + s * BS + b, v);
                }
            }
        } else {
            for (int e = tid; e < H * NC; e += NTT) {
                const int r = e / NC, c = e - r * NC;
                const int s = c / BS, b = c - s * BS;
                const int t = t0n + s;
                float v = 0.f;
                if (t >= 0) {
                    float acc = ibias[r];
                    for (int j = 0; j < K; ++j) {
                        const int tau = t - (K - 1) + j;
                        if (tau < 0) continue;
                        const int n = tau - P.R - 1;
                        const float av = (n >= -SEG) ? hs[b * B3_HIST + ((n + 64 * B3_HIST) % B3_HIST)] : (is_c ? 0.f : __int_as_float(P.Q / 2));
                        if (is_c) acc += itab[(j * 2) * H + r] + itab[(j * 2 + 1) * H + r] * av;
                        else acc += __ldg(P.in_tab + ((size_t)j * P.Q + __float_as_int(av)) * H + r);
                    }
                    v = softsignf_(acc);
                }
                if (P.umma) {
                    const int xo = b3_uoff(c, r, P.NP);
                    const float hh = b3_tf32_hi(v);
                    __stcg(dst + xo, hh); __stcg(dst + H * P.NP + xo, v - hh);
                } else __stcg(dst + e, v);
            }
        }
        const int it = itn - 1;
        if (tid == 0 && itn > 0) B3_PROF(5);
        tsync();
        if (tid == 0 && itn > 0) B3_PROF(6);
        if (tid == 0) b3_signal(P.cnt + (size_t)m * B3_CNT_STRIDE);
    };

    if (P.build_x) {
        for (int e = tid; e < nown * BS * B3_HIST; e += NTT) {
            const int ml = e / (BS * B3_HIST), rest = e - ml * (BS * B3_HIST);
            __stcg(P.hist_g + (size_t)(hc + ml * NVT) * BS * B3_HIST + rest, hist[e]);
        }
    }
    for (int i = 0; i < nown; ++i) input_stage(0, hc + i * NVT, i);

    const unsigned int* cnt_in = P.cnt + (size_t)P.tail_cnt * M * B3_CNT_STRIDE;
    for (int it = 0; it < P.n_it; ++it) {
        const bool active = it >= P.it0;
        const int istep = it - P.it0;
        for (int ml = 0; ml < nown; ++ml) {
            const int m = hc + ml * NVT;
            if (istep >= __ldg(P.slot_nit + m)) continue;          // every utterance of the slot is complete
            // ---- the noise of this step does not depend on the network: draw it while the head is still busy
            if (active) {
                if (is_c) {
                    for (int e = tid; e < BS * SEG; e += NTT) {
                        const int b = e / SEG, j = e - b * SEG;
                        int bglob = m * BS + b; if (bglob >= P.B) bglob = P.B - 1;
                        const long long n = (long long)istep * SEG + j;
                        const float ev = P.noise ? P.noise[(size_t)bglob * P.n_out + n] : laplace_eps(P.seed, bglob, n);
                        const float sgn = (ev > 0.f) ? 1.f : ((ev < 0.f) ? -1.f : 0.f);
                        nz[e] = sgn * log1pf(-2.f * fabsf(ev));      // sign(eps)*log1p(-2|eps|) of cswnv_shift1.py:374
                    }
                } else {
                    for (int e = tid; e < BS * n2; e += NTT) {
                        const int b = e / n2, rr = e - b * n2;
                        int bglob = m * BS + b; if (bglob >= P.B) bglob = P.B - 1;
                        const float qv = P.noise ? P.noise[((size_t)bglob * P.n_out + istep) * P.Q + rr] : softmax_q(P.seed, bglob, istep, rr);
                        nz[e] = -logf(qv);
                    }
                }
            }
            // ---- wait for the head (flow control during warm-up, data when active)
            if (tid == 0)
                *flag = b3_cnt_wait(cnt_in + (size_t)m * B3_CNT_STRIDE, (unsigned)P.tail_nprod * (unsigned)(it + 1), P, 200) ? 1 : 0;
            tsync();
            if (*flag == 0) return;
            if (tid == 0) B3_PROF(0);
            if (active) {
                const float* src = P.tiles + P.tail_in_off + (size_t)m * P.tail_slot_f;      // [rows][BS]
                if (P.tail_gemm) {
                    for (int e = tid; e < n1 * BS; e += NTT) {
                        const int k = e / BS, b = e - k * BS;
                        x2[b * (n1p + 4) + k] = __ldcg(src + e);
                    }
                    tsync();
                    if (tid == 0) B3_PROF(7);
                    // out_2: few rows; 8 lanes per (row, utterance), each an interleaved eighth of k
                    for (int e0 = 0; e0 < n2 * BS; e0 += NTT / 8) {
                        const int e = e0 + (tid >> 3), kq = tid & 7;
                        float v = 0.f;
                        int rr = 0, b = 0;
                        if (e < n2 * BS) {
                            rr = e / BS; b = e - rr * BS;
                            const float4* w4 = reinterpret_cast<const float4*>(w2 + (size_t)rr * n1p);
                            const float4* x4 = reinterpret_cast<const float4*>(x2 + (size_t)b * (n1p + 4));   // +4: the utterances of a warp hit different banks
                            float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
                            for (int c = kq; c < (n1p >> 2); c += 8) {
                                const float4 w = w4[c], x = x4[c];
                                a0 = fmaf(w.x, x.x, fmaf(w.y, x.y, a0));
                                a1 = fmaf(w.z, x.z, fmaf(w.w, x.w, a1));
                            }
                            v = a0 + a1;
                        }
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 2);
                        v += __shfl_xor_sync(0xffffffffu, v, 1);
                        if (kq == 0 && e < n2 * BS) oc[b * n2p + rr] = v + b2s[rr];
                    }
                } else {
                    for (int e = tid; e < n2 * BS; e += NTT) {
                        const int rr = e / BS, b = e - rr * BS;
                        oc[b * n2p + rr] = __ldcg(src + e);
                    }
                }
                tsync();
                if (tid == 0) B3_PROF(1);
                float* hs = hist + (size_t)ml * BS * B3_HIST;
                if (is_c) {
                    // Laplace / LPC sampler (cswnv_shift1.py:368-391).  Everything but the LP feedback is evaluated by one
                    // thread per sample: loc - scale*sign(eps)*log1p(-2|eps|); the recursion over the segment is then short
                    float* sb = oc + BS * n2p;                             // [BS][SEG]
                    for (int e = tid; e < BS * SEG; e += NTT) {
                        const int b = e / SEG, j = e - b * SEG;
                        const float* o = oc + b * n2p;
                        sb[e] = o[j] - sigmoidf_(o[SEG + j]) * nz[e];      // scale = exp(logsigmoid(.))
                    }
                    tsync();
                    if (tid < BS) {
                        const int b = tid;
                        int bglob = m * BS + b;
                        const bool real = bglob < P.B;
                        if (!real) bglob = P.B - 1;
                        const long long nbase = (long long)istep * SEG;
                        const int lpc = P.lpc;
                        const float* o = oc + b * n2p;
                        float* lp = lpb + ((size_t)ml * BS + b) * 8;
                        const float* forced = P.forced ? (const float*)P.forced + (size_t)bglob * P.n_out + nbase : nullptr;
                        float l[8], ac[8], smp[SEG];
#pragma unroll
                        for (int mm = 0; mm < 8; ++mm) {
                            l[mm] = mm < lpc ? lp[mm] : 0.f;
                            ac[mm] = mm < lpc ? o[2 * SEG + lpc - 1 - mm] : 0.f;   // a.flip(-1): a[0] multiplies the newest sample
                        }
#pragma unroll
                        for (int j = 0; j < SEG; ++j) {
                            float lpv = 0.f;
#pragma unroll
                            for (int mm = 0; mm < 8; ++mm) lpv += ac[mm] * l[mm];
                            const float v = fminf(fmaxf(lpv + sb[b * SEG + j], -1.f), 1.f);
                            smp[j] = v;
#pragma unroll
                            for (int mm = 0; mm < 7; ++mm) l[mm] = (mm + 1 < lpc) ? l[mm + 1] : l[mm];
#pragma unroll
                            for (int mm = 0; mm < 8; ++mm)
                                if (mm == lpc - 1) l[mm] = v;
                        }
                        if (forced) {
                            // teacher forcing: inside the segment the recursion ran on its own samples (as the reference's
                            // sampler would); the history handed to the next step is the forced waveform
                            for (int j = 0; j < SEG; ++j) {
                                for (int mm = 0; mm + 1 < lpc; ++mm) lp[mm] = lp[mm + 1];
                                if (lpc > 0) lp[lpc - 1] = forced[j];
                            }
                        } else {
#pragma unroll
                            for (int mm = 0; mm < 8; ++mm)
                                if (mm < lpc) lp[mm] = l[mm];
                        }
#pragma unroll
                        for (int j = 0; j < SEG; ++j) {
                            const float fed = forced ? forced[j] : smp[j];
                            hs[b * B3_HIST + (int)((nbase + j) % B3_HIST)] = fed;
                            if (P.build_x) __stcg(P.hist_g + ((size_t)m * BS + b) * B3_HIST + (int)((nbase + j) % B3_HIST), fed);
                            if (real) ((float*)P.out)[(size_t)bglob * P.n_out + nbase + j] = smp[j];
                        }
                    }
                    if (P.head_out) {
                        for (int e = tid; e < BS * n2; e += NTT) {
                            const int b = e / n2, rr = e - b * n2;
                            const int bglob = m * BS + b;
                            if (bglob < P.B) P.head_out[((size_t)bglob * P.n_iter + istep) * n2 + rr] = oc[b * n2p + rr];
                        }
                    }
                } else {
                    // Gumbel-max: argmax_k(logit_k - log q_k)  (dswnv.py:363-369)
                    for (int b = warp; b < BS; b += (NTT / 32)) {
                        int bglob = m * BS + b;
                        const bool real = bglob < P.B;
                        if (!real) bglob = P.B - 1;
                        float best = -CUDART_INF_F; int bi = 0x7fffffff;
                        for (int rr = lane; rr < n2; rr += 32) {
                            const float o = oc[b * n2p + rr];
                            if (P.head_out && real) P.head_out[((size_t)bglob * P.n_iter + istep) * n2 + rr] = o;
                            const float g = o + nz[b * n2 + rr];
                            if (g > best || (g == best && rr < bi)) { best = g; bi = rr; }
                        }
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1) {
                            const float ob = __shfl_xor_sync(0xffffffffu, best, o2);
                            const int oi = __shfl_xor_sync(0xffffffffu, bi, o2);
                            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                        }
                        if (lane == 0) {
                            int fed = bi;
                            if (real) {
                                ((long long*)P.out)[(size_t)bglob * P.n_out + istep] = bi;
                                if (P.forced) fed = clamp_idx(((const long long*)P.forced)[(size_t)bglob * P.n_out + istep], P.Q);
                            }
                            hs[b * B3_HIST + (istep % B3_HIST)] = __int_as_float(fed);
                        }
                    }
                }
                tsync();
                if (tid == 0) B3_PROF(3);
            }
            if (it + 1 < P.n_it) input_stage(it + 1, m, ml);
            if (tid == 0) B3_PROF(4);
        }
    }
}

// ------------------------------------------------------------------ kernel
template <int SEG, int BS>
__global__ void __launch_bounds__(SWN_NTHREADS, 1) swn_batch_kernel(const __grid_constant__ BatchParams P) {
    extern __shared__ __align__(16) float smem[];
    const BCta& D = P.desc[blockIdx.x];
    if (D.role == ROLE_LAYER) b3_stage<SEG, BS, 0>(P, D, smem);
    else if (D.role == ROLE_DENSE) b3_stage<SEG, BS, 1>(P, D, smem);
    else if (D.role == ROLE_HEAD) {
        if (P.tail_inst == 2) b3_tail<SEG, BS, SWN_NTHREADS / 2>(P, D, smem, threadIdx.x / (SWN_NTHREADS / 2), 2);
        else b3_tail<SEG, BS, SWN_NTHREADS>(P, D, smem, 0, 1);
    }
}

}  // namespace swn
