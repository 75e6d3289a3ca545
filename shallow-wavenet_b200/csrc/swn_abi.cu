// C ABI (include/swn_b200.h) of the B200-native shallow-wavenet generation path:
// model handle, state_dict ingestion + weight folding/packing, CTA work plan, conditioning
// and the launch of the persistent generation kernel.  No CPU compute path exists here:
// every entry point that produces samples launches the sm_100a kernels.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/swn_b200.h"
#include "swn_gen.cuh"
#include "swn_core.cuh"
#include "swn_batch.cuh"
#include "swn_launch.h"
#include "swn_fwd.h"
#include "swn_post.h"

namespace {

thread_local char g_err[1024] = "";

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define CUDA_OK(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) return fail("%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int ensure(size_t count) {
        if (count <= n && p) return 0;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        n = count;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

struct swn_model {
    swn_config cfg;
    int device = 0;
    int n_sm = 0;
    bool finalized = false;
    std::map<std::string, HostTensor> t;

    // derived
    int H = 0, Hp = 0, S = 0, Sp = 0, K = 0, L = 0, seg = 1, lpc = 0, n1 = 0, n1p = 0, n2 = 0, Q = 0, U = 0, R = 0;
    int Ca = 0, Cap = 0, NP = 0, NCH = 0;
    std::vector<int> dil;
    int NH = 4, G = 0, Geff = 0;

    // packed host weights + per-CTA plan pieces that do not depend on the batch
    std::vector<float> wbuf_h;
    struct LayerCta { int pair0, npairs, skip0, nskip; int64_t w_off, b_off, ws_off; };
    struct HeadCta { int a0, na, b0, nb, c0, nc, i0, ni; int64_t wa_off, ca_off, wb_off, bb_off, wc_off, bc_off; };
    std::vector<std::vector<LayerCta>> lcta;   // [L][Geff]
    std::vector<HeadCta> hcta;                 // [NH]

    // device: weights and tables
    DevBuf<float> wbuf, pconst, gain, in_tab, in_bias, xa_tab;
    // conditioning weights
    struct ConvW { DevBuf<float> w_t, b; int Ci, Co, ks, dil, pad; };
    std::vector<ConvW> aux_convs;              // scale_in followed by the conv_aux layers
    DevBuf<float> wp;                          // [Cap][NP]
    // conditioning state
    DevBuf<float> act0, act1, ptab;
    int condB = 0, condTf = 0;
    long long plan_key = -1;
    // generation scratch
    DevBuf<CtaDesc> desc;
    DevBuf<unsigned long long> ll;
    DevBuf<float> ring_scratch;
    DevBuf<int32_t> status;
    DevBuf<float> e2e_aux;
    DevBuf<char> e2e_out;
    DevBuf<char> tf_out;        // swn_teacher_forced: where the discarded samples go
    // ---- v2: core cluster + helpers (swn_core.cuh)
    bool core_ok = false;
    int n_core_taps = 0;
    std::vector<int> tap_base, ncore;
    struct HelpLayer { int ntaps = 0, dmin = 0; std::vector<int> delay, tapj; };
    std::vector<HelpLayer> hl;
    std::vector<float> in_tab_h;
    std::vector<float> wbuf2_h;                 // core section, then (per plan) the helper section
    size_t wbuf2_core_floats = 0;
    std::vector<swn::CoreCta> ccta;
    DevBuf<float> wbuf2, ring2;
    DevBuf<swn::CoreCta> cdesc;
    DevBuf<swn::HelpCta> hdesc;
    DevBuf<unsigned long long> ll2;
    long long core_plan_key = -1;
    int core_multi1_ok[2] = {-1, -1};  // the same with one utterance per cluster, for 2 / 3 utterances
    int core_multi_ok = -1;            // side-by-side core clusters of the one-sample-per-step kernel: -1 untried, 0 do not fit, 1 ok
    int core_grid = 0, core_n_help = 0;
    size_t core_smem = 0, core_ring_total = 0;
    swn::CoreParams core_tmpl;
    size_t core_n_hin = 0, core_n_hacc = 0;
    // ---- v3: pipelined batch kernel (swn_batch.cuh)
    struct L3Cta { int pair0, npairs, skip0, nskip, rows, RP; int64_t w_off, b_off, wr_off; };
    struct D3Cta { int64_t w_off, b_off; };
    std::vector<D3Cta> dense3;
    std::vector<float> wbuf3_h;
    size_t wbuf3_base = 0;
    long long dense3_key = -1;
    std::vector<std::vector<L3Cta>> l3;        // [L][g3]
    int g3 = 0, Hp3 = 0;
    int64_t w2_off3 = 0, b2_off3 = 0, cs_off3 = 0;
    DevBuf<float> wbuf3, tiles3, skip3, ring3, hist3;
    DevBuf<unsigned int> cnt3;
    DevBuf<int32_t> slot3;
    std::vector<int32_t> slot3_h;
    DevBuf<swn::BCta> desc3;
    long long plan3_key = -1;
    int grid3 = 0, nh3 = 0;
    size_t smem3 = 0, ring3_total = 0;
    // ---- time-parallel teacher-forced forward (swn_fwd.cu)
    bool fwd_ready = false;
    struct FwdW { DevBuf<float> A, bias; int n_mt = 0, n_kb = 0, C = 0, Cpad = 0, ntaps = 0, rows = 0; std::vector<int> delay; };
    std::vector<FwdW> fwd_gate;                // [L]
    FwdW fwd_skip, fwd_out1, fwd_out2;
    DevBuf<float> fwd_h_hi, fwd_h_lo, fwd_p, fwd_x1_hi, fwd_x1_lo, fwd_x2_hi, fwd_x2_lo;
    DevBuf<int32_t> fwd_idx, fwd_bad;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool cond_timed = false, gen_timed = false;
    swn_info info;

    const HostTensor* get(const std::string& k) const {
        auto it = t.find(k);
        return it == t.end() ? nullptr : &it->second;
    }
};

namespace {

bool shape_is(const HostTensor* x, std::initializer_list<int64_t> s) {
    return x && x->shape == std::vector<int64_t>(s);
}

int need(const swn_model* m, const std::string& key, std::initializer_list<int64_t> s, const HostTensor** out) {
    const HostTensor* x = m->get(key);
    if (!x) return fail("state_dict entry '%s' was not supplied", key.c_str());
    if (!shape_is(x, s)) {
        std::string got, want;
        for (auto v : x->shape) got += std::to_string(v) + ",";
        for (auto v : s) want += std::to_string(v) + ",";
        return fail("state_dict entry '%s' has shape (%s) but the constructor arguments imply (%s)", key.c_str(),
                    got.c_str(), want.c_str());
    }
    *out = x;
    return 0;
}

int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

template <typename T>
int upload(DevBuf<T>& d, const std::vector<T>& h) {
    if (d.ensure(h.size())) return 1;
    if (!h.empty()) CUDA_OK(cudaMemcpy(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

void split_range(int n, int parts, int i, int* lo, int* cnt) {
    const int a = (int)((long long)n * i / parts), b = (int)((long long)n * (i + 1) / parts);
    *lo = a; *cnt = b - a;
}

}  // namespace


// ---- v1 packing: L layer groups of `geff` CTAs + NH head CTAs (weights of a head group are shared by all head groups)
static int pack_v1(swn_model* m, int geff) {
    const swn_config& c = m->cfg;
    const bool is_c = c.kind == SWN_KIND_CSWNV;
    const int H = m->H, Hp = m->Hp, S = m->S, Sp = m->Sp, K = m->K, L = m->L;
    const int n1 = m->n1, n1p = m->n1p, n2 = m->n2, twoH = 2 * H;
    const HostTensor *x, *y;
    m->Geff = geff;
    // ---------------- in-loop weights, packed per CTA
    std::vector<float>& W = m->wbuf_h;
    W.clear();
    auto alloc = [&](size_t n) { size_t off = (W.size() + 3) & ~size_t(3); W.resize(off + n, 0.f); return (int64_t)off; };
    m->lcta.assign(L, {});
    std::vector<const HostTensor*> skw(L), skb(L);
    for (int l = 0; l < L; ++l) {
        if (need(m, "out_skip." + std::to_string(l) + ".weight", {S, H, 1}, &skw[l]) ||
            need(m, "out_skip." + std::to_string(l) + ".bias", {S}, &skb[l])) return 1;
    }
    for (int l = 0; l < L; ++l) {
        const std::string base = "dil_h." + std::to_string(l) + ".conv";
        if (need(m, base + ".weight", {twoH, H, K}, &x) || need(m, base + ".bias", {twoH}, &y)) return 1;
        for (int g = 0; g < geff; ++g) {
            swn_model::LayerCta lc;
            split_range(H, geff, g, &lc.pair0, &lc.npairs);
            split_range(S, geff, g, &lc.skip0, &lc.nskip);
            if (l == 0) lc.nskip = 0;
            const int np = lc.npairs;
            lc.w_off = alloc((size_t)K * np * 2 * Hp);
            for (int j = 0; j < K; ++j)
                for (int pi = 0; pi < np; ++pi)
                    for (int zc = 0; zc < 2; ++zc) {
                        const int row = zc * H + lc.pair0 + pi;
                        float* dst = W.data() + lc.w_off + ((size_t)(j * np + pi) * 2 + zc) * Hp;
                        for (int k = 0; k < H; ++k) dst[k] = x->data[((size_t)row * H + k) * K + j];
                    }
            lc.b_off = alloc((size_t)np * 2);
            for (int pi = 0; pi < np; ++pi) {
                W[lc.b_off + pi * 2] = y->data[lc.pair0 + pi];
                W[lc.b_off + pi * 2 + 1] = y->data[H + lc.pair0 + pi];
            }
            lc.ws_off = alloc((size_t)std::max(lc.nskip, 1) * Hp);
            for (int r = 0; r < lc.nskip; ++r)
                for (int k = 0; k < H; ++k)
                    W[lc.ws_off + (size_t)r * Hp + k] = skw[l - 1]->data[(size_t)(lc.skip0 + r) * H + k];
            m->lcta[l].push_back(lc);
        }
    }
    {
        const HostTensor *w1, *b1, *w2, *b2;
        if (need(m, "out_1.weight", {n1, S, 1}, &w1) || need(m, "out_1.bias", {n1}, &b1) ||
            need(m, "out_2.weight", {n2, n1, 1}, &w2) || need(m, "out_2.bias", {n2}, &b2)) return 1;
        m->hcta.clear();
        for (int hc = 0; hc < m->NH; ++hc) {
            swn_model::HeadCta q;
            split_range(S, m->NH, hc, &q.a0, &q.na);
            split_range(n1, m->NH, hc, &q.b0, &q.nb);
            if (is_c) { q.c0 = 0; q.nc = n2; } else split_range(n2, m->NH, hc, &q.c0, &q.nc);
            split_range(H, m->NH, hc, &q.i0, &q.ni);
            q.wa_off = alloc((size_t)std::max(q.na, 1) * Hp);
            q.ca_off = alloc((size_t)std::max(q.na, 1));
            for (int r = 0; r < q.na; ++r) {
                for (int k = 0; k < H; ++k) W[q.wa_off + (size_t)r * Hp + k] = skw[L - 1]->data[(size_t)(q.a0 + r) * H + k];
                double cs = 0.0;
                for (int l = 0; l < L; ++l) cs += skb[l]->data[q.a0 + r];
                W[q.ca_off + r] = (float)cs;
            }
            q.wb_off = alloc((size_t)std::max(q.nb, 1) * Sp);
            q.bb_off = alloc((size_t)std::max(q.nb, 1));
            for (int r = 0; r < q.nb; ++r) {
                for (int k = 0; k < S; ++k) W[q.wb_off + (size_t)r * Sp + k] = w1->data[(size_t)(q.b0 + r) * S + k];
                W[q.bb_off + r] = b1->data[q.b0 + r];
            }
            q.wc_off = alloc((size_t)std::max(q.nc, 1) * n1p);
            q.bc_off = alloc((size_t)std::max(q.nc, 1));
            for (int r = 0; r < q.nc; ++r) {
                for (int k = 0; k < n1; ++k) W[q.wc_off + (size_t)r * n1p + k] = w2->data[(size_t)(q.c0 + r) * n1 + k];
                W[q.bc_off + r] = b2->data[q.c0 + r];
            }
            m->hcta.push_back(q);
        }
    }

    if (upload(m->wbuf, m->wbuf_h)) return 1;
    m->info.weight_bytes = (int64_t)m->wbuf_h.size() * 4;
    m->plan_key = -1;
    return 0;
}

static swn::core_fns find_core_fns(int nch, int BS, bool multi) {
    switch (nch) {
        case 1: return swn::find_core_kernel_nch1(BS, multi);
        case 2: return swn::find_core_kernel_nch2(BS, multi);
        case 3: return swn::find_core_kernel_nch3(BS, multi);
        case 4: return swn::find_core_kernel_nch4(BS, multi);
    }
    return swn::core_fns{nullptr, nullptr};
}

static cudaError_t m_core_launch_dispatch(swn_model* m, int BS, bool multi, int grid, const swn::CoreParams& P, cudaStream_t st) {
    swn::core_fns fn = find_core_fns(m->NCH, BS, multi);
    if (!fn.launch) return cudaErrorInvalidValue;
    int dummy = 0;
    cudaError_t e = fn.prepare(m->core_smem, &dummy);
    if (e != cudaSuccess) return e;
    return fn.launch(P, grid, m->core_smem, st);
}

// ---- v2 launch: returns 0 = launched, 2 = not applicable (caller falls back to v1), 1 = error
// b0 / first / last: this launch handles utterances b0 .. b0+B-1 of the conditioned batch (swn_generate cuts batches of
// 3-7 utterances into launches of two, see there); pointers are offset here, the Philox counter in the kernel.
// ncl_plan > 1 (swn_core_kernel2<.., MULTI>): up to ncl_plan core clusters side by side, each with up to BS utterances of its
// own, sharing the helper CTAs (max_cl - ncl_plan helper clusters; per-cluster hand-off rings and partial sums); this launch
// uses ceil(B / BS) of them.
static int generate_core(swn_model* m, const swn_gen_args* a, cudaStream_t st, int B, int Tf, int64_t n_iter,
                         int b0 = 0, bool first = true, bool last = true, int force_bs = 0, int ncl_plan = 1) {
    const int H = m->H, Hp = m->Hp, K = m->K, L = m->L, S = m->S, Q = m->Q;
    const bool is_c = m->cfg.kind == SWN_KIND_CSWNV;
    if (m->seg != 1) return 2;                           // (seg > 1: generate_coreseg)
    int BS = 0;
    for (int t : {1, 2, 4}) if (!BS && t >= B) BS = t;
    if (a->batch_tile > 0) BS = a->batch_tile;
    if (force_bs > 0) BS = force_bs;                     // (an odd last launch keeps the plan of the others)
    if (BS != 1 && BS != 2 && BS != 4) return 2;
    const bool multi = ncl_plan > 1;
    const int ncl = multi ? (B + BS - 1) / BS : 1;         // core clusters of this launch
    if (ncl_plan < 1 || ncl_plan > swn::CORE_MAXCL || ncl > ncl_plan) return 2;
    if (B > BS * ncl) return 2;
    if ((K - 1) * BS > 32) return 2;
    const int NC = BS;
    swn::core_fns fn = find_core_fns(m->NCH, BS, multi);
    if (!fn.launch) return 2;
    const long long key = BS + 10 * (ncl_plan - 1);
    if (key != m->core_plan_key) {
        CUDA_OK(cudaStreamSynchronize(st));
        const int smem_floats = SWN_MAX_SMEM / 4;
        // ---- core CTA shared-memory layout: IDENTICAL in all 16 CTAs (peers address each other's buffers
        // with mapa on their own offsets), so it is sized for the largest row counts
        size_t core_used = 0;
        {
            int mx_skip = 0, mx_nb = 0, mx_nc = 0, mx_ni = 0;
            for (auto& q : m->ccta) {
                mx_skip = std::max(mx_skip, q.nskip); mx_nb = std::max(mx_nb, q.nb);
                mx_nc = std::max(mx_nc, q.nc); mx_ni = std::max(mx_ni, q.ni);
            }
            size_t u = 0;
            auto take = [&](size_t n) { size_t o = u; u += (n + 3) & ~size_t(3); return (int32_t)o; };
            swn::CoreCta lay;
            lay.sm_skip = take((size_t)L * mx_skip * Hp);
            lay.sm_w1 = take((size_t)mx_nb * m->Sp);
            lay.sm_w2 = take((size_t)mx_nc * m->n1p);
            lay.sm_tin = take(is_c ? 0 : (size_t)Q * ((mx_ni + 3) & ~3));
            lay.sm_ring = take((size_t)L * 2 * 16 * 2 * BS);
            lay.sm_hx = take((size_t)(L + 1) * NC * Hp);            // hx, x1, x2, gx stay contiguous (zeroed as one block)
            lay.sm_x1 = take((size_t)NC * m->Sp);
            lay.sm_x2 = take((size_t)NC * m->n1p);
            lay.sm_gx = take((size_t)NC * 256);
            lay.sm_pre = take((size_t)2 * L * 16 * NC * (sizeof(swn::PreSlot) / 4));
            lay.sm_misc = take((size_t)NC * (16 + 8 + 16) + (size_t)L * 16 * 4 + (size_t)(8 * 16 + 2 * 8 * 16) * NC + 8 * 2 * 16 + (size_t)2 * 16 * NC + (size_t)2 * L * 2 * 16 * NC * 2);   // ... + Gumbel noise [2][16][NC] + helper words [2][L][2][16*NC] (u64)
            lay.sm_tmem_slot = take(4 + 2 * (size_t)((L + 4) * BS + 4));  // TMEM base slot, one exchange mbarrier per (vector, column): (L+4) x BS
            core_used = u;
            for (auto& q : m->ccta) {
                q.sm_skip = lay.sm_skip; q.sm_w1 = lay.sm_w1; q.sm_w2 = lay.sm_w2; q.sm_tin = lay.sm_tin;
                q.sm_ring = lay.sm_ring; q.sm_hx = lay.sm_hx; q.sm_x1 = lay.sm_x1; q.sm_x2 = lay.sm_x2;
                q.sm_gx = lay.sm_gx; q.sm_pre = lay.sm_pre; q.sm_misc = lay.sm_misc; q.sm_tmem_slot = lay.sm_tmem_slot;
            }
        }
        if (core_used > (size_t)smem_floats) return 2;
        // ---- how many clusters can be co-resident with the largest footprint we may need
        int max_cl = 0;
        CUDA_OK(fn.prepare(SWN_MAX_SMEM, &max_cl));
        int n_help_layers = 0;
        for (int l = 0; l < L; ++l) n_help_layers += m->hl[l].ntaps > 0;
        if (max_cl < ncl_plan || (n_help_layers > 0 && max_cl < ncl_plan + 1)) return 2;
        const int n_help = n_help_layers ? (max_cl - ncl_plan) * swn::CORE_CL : 0;
        const int GH = n_help_layers ? n_help / n_help_layers : 0;
        if (n_help_layers && GH < 1) return 2;
        // MULTI: CTAs per layer by shared-memory need first (weights + ncl_plan rings of exact length per pair), then the spare
        // ones to the layer with the most pairs x taps per CTA (as generate_coreseg)
        std::vector<int> GHl(L, GH), rlen(L, 0);
        for (int l = 0; l < L; ++l) {
            const auto& hl = m->hl[l];
            if (!hl.ntaps) { GHl[l] = 0; continue; }
            rlen[l] = 4; while (rlen[l] < hl.delay.back() + 2) rlen[l] *= 2;       // power of two
            GHl[l] = std::min(GH, H);
        }
        if (multi) {
            int used_ctas = 0;
            for (int l = 0; l < L; ++l) {
                const auto& hl = m->hl[l];
                if (!hl.ntaps) continue;
                rlen[l] = hl.delay.back() + 2;
                const size_t per_pair = (size_t)hl.ntaps * 2 * Hp + (size_t)ncl_plan * rlen[l] * 2 * BS + 8;
                const size_t room = (size_t)smem_floats - (size_t)ncl_plan * 2 * NC * Hp - 16;
                const int np_max = (int)std::min<size_t>(room / per_pair, (size_t)H);
                if (np_max < 1) return 2;
                GHl[l] = (H + np_max - 1) / np_max;
                used_ctas += GHl[l];
            }
            if (used_ctas > n_help) return 2;
            auto cost = [&](int l) { const int np = (H + GHl[l] - 1) / GHl[l]; return (long long)np * (2 + m->hl[l].ntaps); };
            for (int spare = n_help - used_ctas; spare > 0; --spare) {
                int best = -1;
                for (int l = 0; l < L; ++l)
                    if (GHl[l] > 0 && GHl[l] < H && (best < 0 || cost(l) > cost(best))) best = l;
                if (best < 0) break;
                GHl[best]++;
            }
        }
        // ---- helper CTAs: pack their taps, place their rings
        std::vector<float>& V = m->wbuf2_h;
        V.resize(m->wbuf2_core_floats);
        auto alloc2 = [&](size_t n) { size_t off = (V.size() + 3) & ~size_t(3); V.resize(off + n, 0.f); return (int64_t)off; };
        std::vector<swn::HelpCta> hd;
        size_t ring_total = 0, help_used = 0;
        const int twoH = 2 * H;
        for (int l = 0; l < L; ++l) {
            const auto& hl = m->hl[l];
            if (!hl.ntaps) continue;
            const HostTensor* x = m->get("dil_h." + std::to_string(l) + ".conv.weight");
            const int GHe = GHl[l];
            for (int g = 0; g < GHe; ++g) {
                swn::HelpCta d;
                memset(&d, 0, sizeof(d));
                d.layer = l; d.dil = m->dil[l]; d.ntaps = hl.ntaps; d.dmin = hl.dmin;
                split_range(H, GHe, g, &d.pair0, &d.npairs);
                d.ring_len = rlen[l];                              // power of two (indexed by mask) unless MULTI (modulo)
                for (int j = 0; j < hl.ntaps; ++j) d.delay[j] = hl.delay[j];
                const int np = d.npairs;
                d.w_off = alloc2((size_t)hl.ntaps * np * 2 * Hp);
                for (int j = 0; j < hl.ntaps; ++j)
                    for (int pi = 0; pi < np; ++pi)
                        for (int zc = 0; zc < 2; ++zc) {
                            const int row = zc * H + d.pair0 + pi;
                            float* dst = V.data() + d.w_off + ((size_t)(j * np + pi) * 2 + zc) * Hp;
                            for (int k = 0; k < H; ++k) dst[k] = x->data[((size_t)row * H + k) * K + hl.tapj[j]];
                        }
                size_t u = 0;
                auto take = [&](size_t n) { size_t o = u; u += (n + 3) & ~size_t(3); return (int32_t)o; };
                d.sm_w = take((size_t)hl.ntaps * np * 2 * Hp);
                d.sm_h = take((size_t)ncl_plan * 2 * NC * Hp);
                const size_t ring_f = (size_t)ncl_plan * d.ring_len * np * 2 * BS;      // one ring per core cluster
                d.sm_ring = (int32_t)u;
                d.ring_in_smem = (u + ring_f <= (size_t)smem_floats);
                if (d.ring_in_smem) u += (ring_f + 3) & ~size_t(3);
                else { d.ring_off = (int64_t)ring_total; ring_total += (ring_f + 3) & ~size_t(3); }
                if (u > (size_t)smem_floats) return 2;
                help_used = std::max(help_used, u);
                hd.push_back(d);
            }
        }
        (void)twoH;
        const int n_help_used = (int)hd.size();
        const int grid = swn::CORE_CL * (ncl_plan + (n_help_used + swn::CORE_CL - 1) / swn::CORE_CL);
        if (grid / swn::CORE_CL > max_cl) return 2;
        m->core_smem = std::max(core_used, help_used) * 4;
        m->core_grid = grid; m->core_n_help = n_help_used; m->core_ring_total = ring_total;
        if (upload(m->wbuf2, V)) return 1;
        if (upload(m->cdesc, m->ccta)) return 1;
        if (hd.empty()) hd.push_back(swn::HelpCta());
        if (upload(m->hdesc, hd)) return 1;
        // ---- exchange rings
        swn::CoreParams& P = m->core_tmpl;
        memset(&P, 0, sizeof(P));
        size_t n_hin = 0, n_hacc = 0;
        for (int l = 0; l < L; ++l) {
            const auto& hl = m->hl[l];
            P.dil[l] = m->dil[l]; P.tap_base[l] = m->tap_base[l];
            P.has_help[l] = hl.ntaps > 0; P.hdmin[l] = hl.dmin;
            int p2 = 4; while (p2 < hl.dmin + 2) p2 *= 2;                    // ring depths are powers of two (index masks)
            P.hslots[l] = p2; P.hacc_nt[l] = p2;
            P.hin_off[l] = (int64_t)n_hin; P.hacc_off[l] = (int64_t)n_hacc;
            if (hl.ntaps) { n_hin += (size_t)P.hslots[l] * NC * H; n_hacc += (size_t)P.hacc_nt[l] * 2 * H * NC; }
        }
        m->core_n_hin = n_hin; m->core_n_hacc = n_hacc;
        if (m->ll2.ensure((size_t)ncl_plan * (n_hin + n_hacc) + 8) || m->ring2.ensure(std::max<size_t>(ring_total, 4))) return 1;
        m->core_plan_key = key;
    }
    swn::CoreParams P = m->core_tmpl;
    P.kind = m->cfg.kind; P.H = H; P.Hp = Hp; P.S = S; P.Sp = m->Sp; P.K = K; P.L = L; P.lpc = m->lpc;
    P.n1 = m->n1; P.n1p = m->n1p; P.n2 = m->n2; P.Q = Q; P.U = m->U; P.R = m->R; P.B = B; P.Tf = Tf;
    P.it0 = m->R; P.n_iter = (int)n_iter; P.n_it = m->R + (int)n_iter;
    P.n_core_taps = m->n_core_taps; P.n_help = m->core_n_help; P.audio_in = 0;
    P.n_out = n_iter;
    P.cdesc = m->cdesc.p; P.hdesc = m->hdesc.p; P.wbuf = m->wbuf2.p;
    P.P = m->ptab.p + (size_t)b0 * Tf * L * 2 * H; P.pconst = m->pconst.p; P.gain = m->gain.p;
    P.in_tab = m->in_tab.p; P.in_bias = m->in_bias.p; P.xa_tab = nullptr;
    P.hin_ll = m->ll2.p; P.hacc_ll = m->ll2.p + m->core_n_hin; P.ring_scratch = m->ring2.p;
    const size_t ll_per_cl = m->core_n_hin + m->core_n_hacc;
    P.n_cl = ncl;
    for (int q = 0; q < ncl; ++q) {
        // cluster q: utterances b0 + q*BS .. (at most BS of them); q = 0 is also what the single-cluster kernel reads
        const int bq = b0 + q * BS;
        const size_t esz = is_c ? sizeof(float) : sizeof(long long), off = (size_t)bq * (size_t)n_iter;
        swn::CoreUtt& u = P.cl[q];
        u.b0 = bq; u.nb = std::min(BS, b0 + B - bq);
        u.P = m->ptab.p + (size_t)bq * Tf * L * 2 * H;
        u.noise = a->noise_dev ? a->noise_dev + off * (is_c ? 1 : Q) : nullptr;
        u.forced = a->forced_dev ? (const char*)a->forced_dev + off * esz : nullptr;
        u.seedw = a->seed_dev ? (const char*)a->seed_dev + (size_t)bq * esz : nullptr;
        u.out = (char*)a->out_dev + off * esz;
        u.head_out = a->head_dev ? a->head_dev + off * m->n2 : nullptr;
        u.hin_ll = m->ll2.p + (size_t)q * ll_per_cl;
        u.hacc_ll = u.hin_ll + m->core_n_hin;
    }
    P.b0 = P.cl[0].b0; P.B = multi ? BS : B;
    P.noise = P.cl[0].noise; P.forced = P.cl[0].forced; P.seedw = P.cl[0].seedw; P.out = P.cl[0].out; P.head_out = P.cl[0].head_out;
    P.seed = a->seed; P.status = m->status.p; P.timeout_cycles = 4000000000ULL;
    P.help_delay_ns = a->debug_helper_delay_ns > 0 ? a->debug_helper_delay_ns : 0;
    P.prof = (unsigned long long*)a->prof_dev; P.prof_start = a->prof_start; P.prof_count = a->prof_dev ? a->prof_count : 0;
    CUDA_OK(cudaMemsetAsync(m->ll2.p, 0, (size_t)ncl * ll_per_cl * sizeof(unsigned long long), st));
    if (m->core_ring_total) CUDA_OK(cudaMemsetAsync(m->ring2.p, 0, m->core_ring_total * sizeof(float), st));
    if (first) {        // a failed launch leaves its status for swn_sync; the launches after it see it and stop at once
        CUDA_OK(cudaMemsetAsync(m->status.p, 0, 8 * sizeof(int32_t), st));
        CUDA_OK(cudaEventRecord(m->ev[2], st));
    }
    cudaError_t e = m_core_launch_dispatch(m, BS, multi, m->core_grid - swn::CORE_CL * (ncl_plan - ncl), P, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
        // the clusters cannot all be co-resident right now (SMs held by other work): not applicable, the caller goes on
        // to the cooperatively launched layer-group kernel
        (void)cudaGetLastError();
        fail("cluster kernel not launched: %s (grid %d, %zu B smem)", cudaGetErrorString(e), m->core_grid, m->core_smem);
        return 2;
    }
    if (e != cudaSuccess) return fail("cluster kernel launch failed: %s (grid %d, %zu B smem)", cudaGetErrorString(e), m->core_grid, m->core_smem);
    if (last) CUDA_OK(cudaEventRecord(m->ev[3], st));
    m->gen_timed = true;
    m->info.kernel_launches++;
    m->info.smem_bytes = (int32_t)m->core_smem;
    m->info.batch_tile = BS;
    m->info.n_slots = first ? ncl : m->info.n_slots + ncl; // core clusters (of up to BS utterances), side by side and over launches
    m->info.grid = m->core_grid;
    m->info.kernel_variant = 2;
    m->info.ctas_per_layer = swn::CORE_CL;          // the core cluster; the other CTAs are helpers
    m->info.head_ctas = m->core_n_help;
    return 0;
}


// ---- seg > 1 cluster kernel (swn_coreseg.cuh), utterance b0 of the conditioned batch: 0 = launched, 2 = not applicable, 1 = error
// ncl utterances (b0 .. b0+ncl-1) run side by side, one core cluster each, and share the helper CTAs; ncl_plan = the
// largest ncl of this swn_generate call (it sizes the helper budget: max_cl - ncl_plan helper clusters).
static int generate_coreseg(swn_model* m, const swn_gen_args* a, cudaStream_t st, int Tf, int64_t n_iter, int b0, bool first,
                            bool last, int ncl, int ncl_plan) {
    const int H = m->H, Hp = m->Hp, K = m->K, L = m->L, S = m->S, seg = m->seg;
    swn::core_fns fn = {nullptr, nullptr};
    switch (m->NCH) {
        case 1: fn = swn::find_coreseg_kernel_nch1(seg); break;
        case 2: fn = swn::find_coreseg_kernel_nch2(seg); break;
        case 3: fn = swn::find_coreseg_kernel_nch3(seg); break;
    }
    if (!fn.launch) return 2;
    if (ncl < 1 || ncl > ncl_plan || ncl_plan > swn::CORE_MAXCL) return 2;
    const long long key = 1000 + seg + 100 * ncl_plan;
    if (key != m->core_plan_key) {
        CUDA_OK(cudaStreamSynchronize(st));
        const int smem_floats = SWN_MAX_SMEM / 4;
        size_t core_used = 0;
        {
            int mx_skip = 0, mx_nb = 0, mx_nc = 0;
            for (auto& q : m->ccta) { mx_skip = std::max(mx_skip, q.nskip); mx_nb = std::max(mx_nb, q.nb); mx_nc = std::max(mx_nc, q.nc); }
            size_t u = 0;
            auto take = [&](size_t n) { size_t o = u; u += (n + 3) & ~size_t(3); return (int32_t)o; };
            swn::CoreCta lay;
            lay.sm_skip = take((size_t)L * mx_skip * Hp);
            lay.sm_w1 = take((size_t)mx_nb * m->Sp);
            lay.sm_w2 = take((size_t)mx_nc * m->n1p);
            lay.sm_hx = take((size_t)2 * (L + 1) * seg * Hp);       // hx (both parities), x1, x2 stay contiguous
            lay.sm_x1 = take((size_t)m->Sp);
            lay.sm_x2 = take((size_t)m->n1p);
            lay.sm_pre = take((size_t)2 * L * 16 * seg * 4);
            lay.sm_misc = take((size_t)(32 + 16 + 16 + 8) + (size_t)L * 16 * 4 + 8 * 2 * 16 + 20 * 4 + (size_t)L * 16);
            lay.sm_tmem_slot = take(4 + 2 * (size_t)(3 * L + 3 + 1));
            core_used = u;
            for (auto& q : m->ccta) {
                q.sm_skip = lay.sm_skip; q.sm_w1 = lay.sm_w1; q.sm_w2 = lay.sm_w2; q.sm_tin = 0; q.sm_ring = 0;
                q.sm_hx = lay.sm_hx; q.sm_x1 = lay.sm_x1; q.sm_x2 = lay.sm_x2; q.sm_gx = 0; q.sm_pre = lay.sm_pre;
                q.sm_misc = lay.sm_misc; q.sm_tmem_slot = lay.sm_tmem_slot;
            }
        }
        if (core_used > (size_t)smem_floats) return 2;
        int max_cl = 0;
        CUDA_OK(fn.prepare(SWN_MAX_SMEM, &max_cl));
        int n_help_layers = 0, n_help_taps = 0;
        for (int l = 0; l < L; ++l) { n_help_layers += m->hl[l].ntaps > 0; n_help_taps += m->hl[l].ntaps; }
        if (max_cl < ncl_plan || (n_help_layers > 0 && max_cl < ncl_plan + 1)) return 2;
        const int n_help = n_help_layers ? (max_cl - ncl_plan) * swn::CORE_CL : 0;
        if (n_help_layers && n_help < n_help_layers) return 2;
        // ---- helper CTAs: a share of the helper SMs per layer in proportion to its taps (at least one)
        std::vector<float>& V = m->wbuf2_h;
        V.resize(m->wbuf2_core_floats);
        auto alloc2 = [&](size_t n) { size_t off = (V.size() + 3) & ~size_t(3); V.resize(off + n, 0.f); return (int64_t)off; };
        std::vector<swn::HelpCta> hd;
        size_t help_used = 0;
        // A helper CTA keeps, per gate pair it owns, the layer's helper taps (ntaps x 2 rows x Hp) and one ring of partial
        // sums per core cluster (ring_len x 2): first the fewest CTAs per layer whose pairs fit the shared memory, then the
        // spare CTAs one by one to the layer whose CTAs have the most work per step.
        std::vector<int> GH(L, 0), rlen(L, 0);
        {
            int used_ctas = 0;
            for (int l = 0; l < L; ++l) {
                const auto& hl = m->hl[l];
                if (!hl.ntaps) continue;
                rlen[l] = hl.delay.back() + 2 * seg + 2;
                const size_t per_pair = (size_t)hl.ntaps * 2 * Hp + (size_t)ncl_plan * rlen[l] * 2 + 8;
                const size_t room = (size_t)smem_floats - (size_t)ncl_plan * seg * Hp - 16;
                const int np_max = (int)std::min<size_t>(room / per_pair, (size_t)H);
                if (np_max < 1) return 2;
                GH[l] = (H + np_max - 1) / np_max;
                used_ctas += GH[l];
            }
            if (used_ctas > n_help) return 2;
            // (issue-bound mat-vecs, measured with the helper stamps of tools/gpu_prof_seghelp.py: a CTA's time per step goes
            // with its pairs x (a fixed part - reductions, ring update - plus the taps' multiply-adds))
            auto cost = [&](int l) { const int np = (H + GH[l] - 1) / GH[l]; return (long long)np * (250 + 130 * m->hl[l].ntaps); };
            for (int spare = n_help - used_ctas; spare > 0; --spare) {
                int best = -1;
                for (int l = 0; l < L; ++l)
                    if (GH[l] > 0 && GH[l] < H && (best < 0 || cost(l) > cost(best))) best = l;
                if (best < 0) break;
                GH[best]++;
            }
        }
        for (int l = 0; l < L; ++l) {
            const auto& hl = m->hl[l];
            if (!hl.ntaps) continue;
            const HostTensor* x = m->get("dil_h." + std::to_string(l) + ".conv.weight");
            const int GHe = GH[l];
            for (int g = 0; g < GHe; ++g) {
                swn::HelpCta d;
                memset(&d, 0, sizeof(d));
                d.layer = l; d.dil = m->dil[l]; d.ntaps = hl.ntaps; d.dmin = hl.dmin;
                split_range(H, GHe, g, &d.pair0, &d.npairs);
                d.ring_len = rlen[l];                              // (need not be a power of two: indexed modulo)
                for (int j = 0; j < hl.ntaps; ++j) d.delay[j] = hl.delay[j];
                const int np = d.npairs;
                d.w_off = alloc2((size_t)hl.ntaps * np * 2 * Hp);
                for (int j = 0; j < hl.ntaps; ++j)
                    for (int pi = 0; pi < np; ++pi)
                        for (int zc = 0; zc < 2; ++zc) {
                            const int row = zc * H + d.pair0 + pi;
                            float* dst = V.data() + d.w_off + ((size_t)(j * np + pi) * 2 + zc) * Hp;
                            for (int k = 0; k < H; ++k) dst[k] = x->data[((size_t)row * H + k) * K + hl.tapj[j]];
                        }
                size_t u = 0;
                auto take = [&](size_t n) { size_t o = u; u += (n + 3) & ~size_t(3); return (int32_t)o; };
                d.sm_w = take((size_t)hl.ntaps * np * 2 * Hp);
                d.sm_h = take((size_t)ncl_plan * seg * Hp);
                d.sm_ring = take((size_t)ncl_plan * d.ring_len * np * 2);        // one ring of partial sums per core cluster
                d.ring_in_smem = 1;
                if (u > (size_t)smem_floats) return 2;
                help_used = std::max(help_used, u);
                hd.push_back(d);
            }
        }
        const int n_help_used = (int)hd.size();
        const int grid = swn::CORE_CL * (ncl_plan + (n_help_used + swn::CORE_CL - 1) / swn::CORE_CL);
        if (grid / swn::CORE_CL > max_cl) return 2;
        m->core_smem = std::max(core_used, help_used) * 4;
        m->core_grid = grid; m->core_n_help = n_help_used; m->core_ring_total = 0;
        if (upload(m->wbuf2, V)) return 1;
        if (upload(m->cdesc, m->ccta)) return 1;
        if (hd.empty()) hd.push_back(swn::HelpCta());
        if (upload(m->hdesc, hd)) return 1;
        swn::CoreParams& P = m->core_tmpl;
        memset(&P, 0, sizeof(P));
        size_t n_hin = 0, n_hacc = 0;
        for (int l = 0; l < L; ++l) {
            const auto& hl = m->hl[l];
            P.dil[l] = m->dil[l]; P.tap_base[l] = m->tap_base[l]; P.ncore[l] = m->ncore[l];
            P.has_help[l] = hl.ntaps > 0; P.hdmin[l] = hl.dmin;
            int p2 = 4; while (p2 < hl.dmin / seg + 4) p2 *= 2;                 // input ring: steps the core may run ahead
            P.hslots[l] = p2;
            p2 = 8; while (p2 < hl.dmin + 3 * seg) p2 *= 2;                     // finished sums: time slots in flight
            P.hacc_nt[l] = p2;
            P.hin_off[l] = (int64_t)n_hin; P.hacc_off[l] = (int64_t)n_hacc;
            if (hl.ntaps) { n_hin += (size_t)P.hslots[l] * seg * H; n_hacc += (size_t)P.hacc_nt[l] * 2 * H; }
        }
        m->core_n_hin = n_hin; m->core_n_hacc = n_hacc;
        if (m->ll2.ensure((size_t)ncl_plan * (n_hin + n_hacc) + 8)) return 1;
        m->core_plan_key = key;
    }
    swn::CoreParams P = m->core_tmpl;
    P.kind = m->cfg.kind; P.H = H; P.Hp = Hp; P.S = S; P.Sp = m->Sp; P.K = K; P.L = L; P.lpc = m->lpc;
    P.n1 = m->n1; P.n1p = m->n1p; P.n2 = m->n2; P.Q = m->Q; P.U = m->U; P.R = m->R; P.B = 1; P.Tf = Tf;
    P.seg = seg;
    const int rem = (m->R + 1) % seg;                      // the time-index contract of the layer-group kernel (swn_generate)
    P.t_first = rem == 0 ? 0 : rem - seg;
    P.it0 = (m->R + 1 - P.t_first) / seg - 1;
    // one utterance per launch: it only runs for its own length (the reference generates every utterance to the longest,
    // cswnv_shift1.py:291-295, and truncates afterwards; the rows beyond an utterance's length are not written here)
    // (utterances that share a launch run to the longer of their lengths)
    int64_t steps_b = n_iter;
    if (a->n_samples_list) {
        steps_b = 1;
        for (int q = 0; q < ncl; ++q)
            steps_b = std::max<int64_t>(steps_b, std::min<int64_t>(n_iter, (a->n_samples_list[b0 + q] + seg - 1) / seg));
    }
    P.n_iter = (int)n_iter; P.n_it = P.it0 + (int)steps_b;
    P.n_core_taps = m->n_core_taps; P.n_help = m->core_n_help;
    {   // (field reused as a flag) per-step cluster barrier also outside the warm-up: some chain warp owns no out_1 row
        int min_nb = 1 << 30;
        for (auto& q : m->ccta) min_nb = std::min(min_nb, (int)q.nb);
        P.audio_in = min_nb < 4 * m->NCH ? 1 : 0;
    }
    P.n_out = n_iter * seg;
    P.cdesc = m->cdesc.p; P.hdesc = m->hdesc.p; P.wbuf = m->wbuf2.p;
    P.P = nullptr; P.pconst = m->pconst.p; P.gain = m->gain.p;
    P.in_tab = m->in_tab.p; P.in_bias = m->in_bias.p; P.xa_tab = nullptr;
    P.hin_ll = nullptr; P.hacc_ll = nullptr; P.ring_scratch = nullptr;
    P.n_cl = ncl; P.b0 = b0;
    const size_t ll_per_cl = m->core_n_hin + m->core_n_hacc;
    for (int q = 0; q < ncl; ++q) {
        swn::CoreUtt& u = P.cl[q];
        const int b = b0 + q;
        const size_t off = (size_t)b * (size_t)P.n_out;
        u.b0 = b;
        u.P = m->ptab.p + (size_t)b * Tf * L * seg * 2 * H;
        u.noise = a->noise_dev ? a->noise_dev + off : nullptr;
        u.forced = a->forced_dev ? (const char*)a->forced_dev + off * sizeof(float) : nullptr;
        u.seedw = a->seed_dev ? (const char*)a->seed_dev + (size_t)b * seg * sizeof(float) : nullptr;
        u.out = (char*)a->out_dev + off * sizeof(float);
        u.head_out = a->head_dev ? a->head_dev + (size_t)b * (size_t)n_iter * m->n2 : nullptr;
        u.hin_ll = m->ll2.p + (size_t)q * ll_per_cl;
        u.hacc_ll = u.hin_ll + m->core_n_hin;
    }
    P.seed = a->seed; P.status = m->status.p; P.timeout_cycles = 4000000000ULL;
    P.help_delay_ns = a->debug_helper_delay_ns > 0 ? a->debug_helper_delay_ns : 0;
    P.prof = (unsigned long long*)a->prof_dev; P.prof_start = a->prof_start; P.prof_count = a->prof_dev ? a->prof_count : 0;
    CUDA_OK(cudaMemsetAsync(m->ll2.p, 0, (size_t)ncl * ll_per_cl * sizeof(unsigned long long), st));
    if (first) {
        CUDA_OK(cudaMemsetAsync(m->status.p, 0, 8 * sizeof(int32_t), st));
        CUDA_OK(cudaEventRecord(m->ev[2], st));
    }
    int max_cl_dummy_ = 0;
    const int grid = m->core_grid - swn::CORE_CL * (ncl_plan - ncl);      // core clusters first, then the helpers
    cudaError_t e = fn.prepare(m->core_smem, &max_cl_dummy_);
    if (e == cudaSuccess) e = fn.launch(P, grid, m->core_smem, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
        (void)cudaGetLastError();
        fail("cluster kernel not launched: %s (grid %d, %zu B smem)", cudaGetErrorString(e), m->core_grid, m->core_smem);
        return first ? 2 : 1;
    }
    if (e != cudaSuccess) return fail("cluster kernel launch failed: %s (grid %d, %zu B smem)", cudaGetErrorString(e), m->core_grid, m->core_smem);
    if (last) CUDA_OK(cudaEventRecord(m->ev[3], st));
    m->gen_timed = true;
    m->info.kernel_launches++;
    m->info.smem_bytes = (int32_t)m->core_smem;
    m->info.batch_tile = ncl_plan;                         // utterances side by side in one launch
    m->info.n_slots = first ? ncl : m->info.n_slots + ncl; // utterances, over launches one after the other
    m->info.grid = m->core_grid;
    m->info.kernel_variant = 2;
    m->info.ctas_per_layer = swn::CORE_CL;
    m->info.head_ctas = m->core_n_help;
    return 0;
}


// ---- v3 packing (batch independent): per gated-cell CTA the transposed GEMM operand Wt[k][row] (rows = all
// taps of the owned gate pairs, then the owned out_skip[l-1] rows); the head matrices are packed per plan
static int pack_v3(swn_model* m, int geff) {
    const int H = m->H, S = m->S, K = m->K, L = m->L, twoH = 2 * H;
    const int Hp = round_up(H, 4);
    std::vector<float>& W = m->wbuf3_h;
    W.clear();
    auto alloc = [&](size_t n) { size_t off = (W.size() + 3) & ~size_t(3); W.resize(off + n, 0.f); return (int64_t)off; };
    std::vector<const HostTensor*> skw(L), skb(L);
    for (int l = 0; l < L; ++l) {
        if (need(m, "out_skip." + std::to_string(l) + ".weight", {S, H, 1}, &skw[l]) ||
            need(m, "out_skip." + std::to_string(l) + ".bias", {S}, &skb[l])) return 1;
    }
    m->l3.assign(L, {});
    const HostTensor *x, *y;
    for (int l = 0; l < L; ++l) {
        const std::string base = "dil_h." + std::to_string(l) + ".conv";
        if (need(m, base + ".weight", {twoH, H, K}, &x) || need(m, base + ".bias", {twoH}, &y)) return 1;
        for (int g = 0; g < geff; ++g) {
            swn_model::L3Cta c;
            split_range(H, geff, g, &c.pair0, &c.npairs);
            split_range(S, geff, g, &c.skip0, &c.nskip);
            if (l == 0) c.nskip = 0;
            const int np = c.npairs;
            c.rows = 2 * K * np + c.nskip;
            c.RP = round_up(c.rows, 8);
            c.w_off = alloc((size_t)Hp * c.RP);
            float* wt = W.data() + c.w_off;
            for (int j = 0; j < K; ++j)
                for (int pi = 0; pi < np; ++pi)
                    for (int zc = 0; zc < 2; ++zc) {
                        const int row = (j * np + pi) * 2 + zc, src = zc * H + c.pair0 + pi;
                        for (int k = 0; k < H; ++k) wt[(size_t)k * c.RP + row] = x->data[((size_t)src * H + k) * K + j];
                    }
            for (int r = 0; r < c.nskip; ++r)
                for (int k = 0; k < H; ++k)
                    wt[(size_t)k * c.RP + 2 * K * np + r] = skw[l - 1]->data[(size_t)(c.skip0 + r) * H + k];
            c.b_off = alloc((size_t)np * 2);
            for (int pi = 0; pi < np; ++pi) {
                W[c.b_off + pi * 2] = y->data[c.pair0 + pi];
                W[c.b_off + pi * 2 + 1] = y->data[H + c.pair0 + pi];
            }
            // the same rows row-major [128][Hp] for the tensor-memory resident form (tcgen05 contraction)
            c.wr_off = 0;
            if (c.RP <= 128) {
                c.wr_off = alloc((size_t)128 * Hp);
                const float* wt2 = W.data() + c.w_off;
                float* wr = W.data() + c.wr_off;
                for (int r = 0; r < c.rows; ++r)
                    for (int k = 0; k < H; ++k) wr[(size_t)r * Hp + k] = wt2[(size_t)k * c.RP + r];
            }
            m->l3[l].push_back(c);
        }
    }
    const HostTensor *w2, *b2;
    if (need(m, "out_2.weight", {m->n2, m->n1, 1}, &w2) || need(m, "out_2.bias", {m->n2}, &b2)) return 1;
    m->w2_off3 = alloc((size_t)m->n2 * m->n1p);
    for (int r = 0; r < m->n2; ++r)
        for (int k = 0; k < m->n1; ++k) W[m->w2_off3 + (size_t)r * m->n1p + k] = w2->data[(size_t)r * m->n1 + k];
    m->b2_off3 = alloc(m->n2);
    for (int r = 0; r < m->n2; ++r) W[m->b2_off3 + r] = b2->data[r];
    m->cs_off3 = alloc(S);
    for (int r = 0; r < S; ++r) {
        double cs = 0.0;
        for (int l = 0; l < L; ++l) cs += skb[l]->data[r];
        W[m->cs_off3 + r] = (float)cs;
    }
    m->wbuf3_base = W.size();
    m->g3 = geff; m->Hp3 = Hp;
    m->plan3_key = -1; m->dense3_key = -1;
    return 0;
}

// ---- v3 launch: returns 0 = launched, 2 = not applicable (caller falls back to v1), 1 = error
static int generate_batch(swn_model* m, const swn_gen_args* a, cudaStream_t st, int B, int Tf, int64_t n_iter) {
    const int H = m->H, K = m->K, L = m->L, S = m->S, seg = m->seg, n1 = m->n1, n2 = m->n2;
    const bool is_c = m->cfg.kind == SWN_KIND_CSWNV;
    if (!is_c && m->cfg.audio_in_flag) return 2;
    const int* tiles = swn::batch3_tiles_for_seg(seg);
    if (!tiles[0]) return 2;
    const int smem_floats = SWN_MAX_SMEM / 4;
    const int n2p = round_up(n2, 4);
    const bool tail_gemm = n2 <= 32;
    // cell 0 can evaluate the causal input stage itself (CSWNV: the table is tiny).  Measured slower on B200: the
    // builder warps compete with the GEMM warps for issue slots, so the tail builds the tile unless asked otherwise
    const bool build_env = getenv("SWN_BUILD_X") != nullptr;
    const bool build_x = is_c && K <= 8 && H % 4 == 0 && (!build_env || atoi(getenv("SWN_BUILD_X")) == 1);
    // dense head stages: {input depth, output rows}
    struct DenseSt { int kdim, rows; };
    std::vector<DenseSt> dst = {{H, S}, {S, n1}};
    if (!tail_gemm) dst.push_back({n1, n2});
    const int ND = (int)dst.size();
    // ---- utterances per slot: explicit, or the largest tile that still leaves >= 8 slots in flight (measured: 64 utterances run faster as 8 slots of 8 than as 16 of 4)
    std::vector<int> cand;
    if (a->batch_tile > 0) {
        for (const int* t = tiles; *t; ++t) if (*t == a->batch_tile) cand.push_back(*t);
        if (cand.empty()) return a->kernel_sel >= 3 ? fail("swn_generate: batch_tile=%d is not instantiated for the batch kernel at seg=%d", a->batch_tile, seg) : 2;
    } else {
        int first = -1;
        for (const int* t = tiles; *t; ++t) if (first < 0 && (B + *t - 1) / *t >= 8) first = *t;
        for (const int* t = tiles; *t; ++t) if (first < 0 ? true : *t <= first) cand.push_back(*t);
        if (first < 0) std::reverse(cand.begin(), cand.end());     // small batches: smallest tile first
    }
    int BS = 0, M = 0, NT = 0, grid = 0, geff = 0, NPsel = 0, ndb_sel = 2, tinst_sel = 1;
    bool umma = false;
    const bool want_umma = a->kernel_sel != 4 && !(getenv("SWN_NO_UMMA") && atoi(getenv("SWN_NO_UMMA")) == 1);
    size_t smem_max = 0, ring_total = 0, tiles_total = 0;
    std::vector<int> nd(ND, 1);
    std::vector<swn::BCta> desc;
    std::vector<int64_t> tile_off;           // per stage input tile offset (stages: L cells, ND dense, tail)
    for (int bs : cand) {
        const int NC = bs * seg, OS = NC | 1, OSd = bs | 1;
        if (((size_t)H * NC * 4) % 16 != 0 || ((size_t)H * bs * 4) % 16 != 0 || ((size_t)S * bs * 4) % 16 != 0 ||
            ((size_t)n1 * bs * 4) % 16 != 0) continue;
        const int Mx = (B + bs - 1) / bs, Bp = Mx * bs;
        // CTAs a dense stage needs so that its rows fit in shared memory
        bool ok = true;
        int nd_min_total = 0;
        for (int d = 0; d < ND && ok; ++d) {
            const int kp = round_up(dst[d].kdim, 4);
            int n = 1;
            for (;; ++n) {
                const int rows = (dst[d].rows + n - 1) / n, RP = round_up(rows, 8);
                const size_t vw = bs % 4 == 0 ? 4 : 1;
                // a single operand buffer is enough to make one CTA hold the whole stage (e.g. out_skip[L-1] at 16 utterances)
                const size_t u = (size_t)kp * RP + (size_t)kp * bs + (size_t)RP * OSd +
                                 ((size_t)rows * bs / vw > 4 * 352 ? (size_t)rows * bs : 0) + 64;
                if (u <= (size_t)smem_floats) break;
                if (n >= dst[d].rows) { ok = false; break; }
            }
            nd[d] = n; nd_min_total += n;
        }
        if (!ok) continue;
        const int G = (m->n_sm - nd_min_total - 1) / L;
        if (G < 1) continue;
        const int ge = std::min(G, H);
        if (ge != m->g3) {
            CUDA_OK(cudaStreamSynchronize(st));
            if (pack_v3(m, ge)) return 1;
        }
        const int Hp = m->Hp3;
        // tensor-core contraction for the gated cells: rows fit one M=128 tile, weights (tf32 hi + lo) and two
        // accumulator tiles fit the 512 columns of tensor memory
        const int NPc = round_up(NC, 16);
        bool um = want_umma && H % 8 == 0 && Hp == H && 2 * H + NPc <= 512 && NPc <= 256;
        const int ndb = 2 * H + 2 * NPc <= 512 ? 2 : 1;        // accumulator tiles that fit beside the weights
        for (int l = 0; l < L && um; ++l)
            for (int g = 0; g < ge; ++g) if (m->l3[l][g].RP > 128) um = false;
        int left = m->n_sm - L * ge - nd_min_total;              // >= 1
        int nt = std::max(1, std::min(std::min(Mx, 2), left));
        left -= nt;
        for (int d = 0; left > 0 && d < 8 * ND; ++d)              // spare SMs: more CTAs per dense stage (up to 8 each)
            if (nd[d % ND] < 8 && nd[d % ND] < dst[d % ND].rows) { ++nd[d % ND]; --left; }
        bool fits = true;
        size_t used_max = 0, rt = 0;
        int n_cta = L * ge + nt;
        for (int d = 0; d < ND; ++d) n_cta += nd[d];
        desc.assign((size_t)n_cta, swn::BCta());
        // tiles: cell inputs [H][NC] x L, last-column tile [H][bs], dense outputs
        tile_off.assign(L + ND + 1, 0);
        size_t tt = 0;
        const size_t cell_tile_f = um ? (size_t)2 * H * NPc : (size_t)H * NC;
        for (int l = 0; l < L; ++l) { tile_off[l] = (int64_t)tt; tt += (size_t)Mx * cell_tile_f; }
        tile_off[L] = (int64_t)tt; tt += (size_t)Mx * H * bs;
        for (int d = 0; d < ND; ++d) { tile_off[L + 1 + d] = (int64_t)tt; tt += (size_t)Mx * dst[d].rows * bs; }
        int cta = 0;
        for (int l = 0; l < L && fits; ++l)
            for (int g = 0; g < ge; ++g, ++cta) {
                const auto& c = m->l3[l][g];
                swn::BCta& d = desc[cta];
                memset(&d, 0, sizeof(d));
                d.role = ROLE_LAYER; d.layer = l; d.hindex = g; d.dil = m->dil[l];
                d.pair0 = c.pair0; d.npairs = c.npairs; d.skip0 = c.skip0; d.nskip = c.nskip; d.rows = c.rows; d.RP = c.RP;
                d.kdim = H; d.kp = Hp;
                d.ring_len = (K - 1) * d.dil + seg;
                d.cnt_in = l; d.cnt_out = l + 1;
                d.nprod_in = l == 0 ? 1 : ge;
                d.is_last = l == L - 1;
                d.in_off = tile_off[l]; d.in_slot_f = (int32_t)cell_tile_f;
                d.out_off = tile_off[l + 1]; d.out_slot_f = d.is_last ? H * bs : (int32_t)cell_tile_f;
                d.w_off = c.w_off; d.b_off = c.b_off; d.wr_off = c.wr_off;
                size_t u = 0;
                auto take = [&](size_t n) { size_t o = u; u += (n + 3) & ~size_t(3); return (int32_t)o; };
                d.sm_w = take(um ? 0 : (size_t)Hp * c.RP);
                // two operand buffers (prefetch of the next slot) unless two large tensor-core tiles do not fit
                d.nxb = (um && 2 * cell_tile_f + (size_t)128 * OS + (size_t)2 * c.npairs * NC * 4 + 64 > (size_t)smem_floats) ? 1 : 2;
                d.sm_x = take(um ? d.nxb * cell_tile_f : (size_t)2 * Hp * NC);
                d.sm_o = take((size_t)(um ? 128 : c.RP) * OS);
                d.sm_pre = take((size_t)2 * c.npairs * NC * 4);
                d.sm_bar = take(48);
                d.sm_build = take((build_x && l == 0) ? (size_t)K * 2 * H + 2 * H + 4 * (size_t)bs * (seg + 7) : 0);
                const size_t ring_f = (size_t)d.ring_len * 2 * c.npairs * Bp;
                d.sm_ring = (int32_t)u;
                d.ring_in_smem = u + ring_f <= (size_t)smem_floats;
                if (d.ring_in_smem) u += (ring_f + 3) & ~size_t(3);
                else { d.ring_off = (int64_t)rt; rt += (ring_f + 3) & ~size_t(3); }
                if (u > (size_t)smem_floats) { fits = false; break; }
                used_max = std::max(used_max, u);
            }
        if (!fits) continue;
        for (int sd = 0; sd < ND && fits; ++sd)
            for (int i = 0; i < nd[sd]; ++i, ++cta) {
                swn::BCta& d = desc[cta];
                memset(&d, 0, sizeof(d));
                d.role = swn::ROLE_DENSE; d.layer = sd; d.hindex = i;
                split_range(dst[sd].rows, nd[sd], i, &d.row0, &d.nrows);
                d.rows = d.nrows; d.RP = round_up(d.nrows, 8);
                d.kdim = dst[sd].kdim; d.kp = round_up(d.kdim, 4);
                d.cnt_in = L + sd; d.cnt_out = L + sd + 1;
                d.nprod_in = sd == 0 ? ge : nd[sd - 1];
                d.relu = sd < 2; d.add_skips = sd == 0;
                d.in_off = tile_off[L + sd]; d.in_slot_f = d.kdim * bs;
                d.out_off = tile_off[L + 1 + sd]; d.out_slot_f = dst[sd].rows * bs;
                size_t u = 0;
                auto take = [&](size_t n) { size_t o = u; u += (n + 3) & ~size_t(3); return (int32_t)o; };
                d.sm_w = take((size_t)d.kp * d.RP);
                {
                    const size_t vw = bs % 4 == 0 ? 4 : 1;
                    const size_t rest = (size_t)d.kp * d.RP + (size_t)d.RP * OSd + ((size_t)d.nrows * bs / vw > 4 * 352 ? (size_t)d.nrows * bs : 0) + 64;
                    d.nxb = rest + 2 * (size_t)d.kp * bs <= (size_t)smem_floats ? 2 : 1;
                }
                d.sm_x = take((size_t)d.nxb * d.kp * bs);
                d.sm_o = take((size_t)d.RP * OSd);
                {   // the epilogue keeps 4 items per thread in registers; shared memory only for what is beyond
                    const size_t vw = bs % 4 == 0 ? 4 : 1, nitem = (size_t)d.nrows * bs / vw;
                    d.sm_pre = take(nitem > 4 * 352 ? (size_t)d.nrows * bs : 0);
                }
                d.sm_bar = take(48);
                if (u > (size_t)smem_floats) { fits = false; break; }
                used_max = std::max(used_max, u);
            }
        if (!fits) continue;
        // two half-size tail instances per tail CTA once the ring is throughput-bound (their short dependent phases hide each other)
        int tinst = Mx >= 16 ? 2 : 1;
        if (getenv("SWN_TAIL_INST")) tinst = atoi(getenv("SWN_TAIL_INST")) == 2 ? 2 : 1;
        const int cta_tail0 = cta;
      retry_tail:
        cta = cta_tail0;
        const int MO = (Mx + nt * tinst - 1) / (nt * tinst);
        for (int hc = 0; hc < nt; ++hc, ++cta) {
            swn::BCta& d = desc[cta];
            memset(&d, 0, sizeof(d));
            d.role = ROLE_HEAD; d.hindex = hc;
            size_t u = 0;
            auto take = [&](size_t n) { size_t o = u; u += (n + 3) & ~size_t(3); return (int32_t)o; };
            d.sm_x2 = take((size_t)bs * (m->n1p + 4));
            d.sm_oc = take((size_t)bs * n2p + (size_t)bs * seg);
            d.sm_w2 = take(tail_gemm ? (size_t)n2 * m->n1p : 0);
            d.sm_tin = take((size_t)H * NC);
            d.sm_itab = take((is_c ? (size_t)K * 2 * H : 0) + 2 * H + n2);
            d.sm_ain = take(std::max((size_t)NC * K * 2, (size_t)bs * ((seg + 7) | 1)));
            d.sm_noise = take(is_c ? (size_t)bs * seg : (size_t)bs * m->Q);
            d.sm_hist = take((size_t)MO * bs * swn::B3_HIST);
            d.sm_lpb = take((size_t)MO * bs * 8);
            d.sm_misc = take(8);
            d.tail_stride = (int32_t)u;
            u *= tinst;
            if (u > (size_t)smem_floats && tinst == 2) { tinst = 1; goto retry_tail; }     // two instances do not fit: one
            if (u > (size_t)smem_floats) { fits = false; break; }
            used_max = std::max(used_max, u);
        }
        if (!fits) continue;
        tinst_sel = tinst;
        BS = bs; M = Mx; NT = nt; grid = cta; geff = ge; smem_max = used_max * 4; ring_total = rt; tiles_total = tt;
        umma = um; NPsel = NPc; ndb_sel = ndb;
        break;
    }
    if (!BS) return 2;
    swn::batch_launch_fn fn = swn::find_batch_kernel(seg, BS);
    if (!fn) return 2;
    // ---- dense stage weights depend on the CTA split: pack them behind the batch-independent part
    long long dkey = ((long long)geff << 40);
    for (int d = 0; d < ND; ++d) dkey = dkey * 16 + nd[d];
    if (dkey != m->dense3_key) {
        CUDA_OK(cudaStreamSynchronize(st));
        std::vector<float>& W = m->wbuf3_h;
        W.resize(m->wbuf3_base);
        auto alloc = [&](size_t n) { size_t off = (W.size() + 3) & ~size_t(3); W.resize(off + n, 0.f); return (int64_t)off; };
        const HostTensor *skw, *w1, *b1, *w2, *b2;
        if (need(m, "out_skip." + std::to_string(L - 1) + ".weight", {S, H, 1}, &skw) ||
            need(m, "out_1.weight", {n1, S, 1}, &w1) || need(m, "out_1.bias", {n1}, &b1) ||
            need(m, "out_2.weight", {n2, n1, 1}, &w2) || need(m, "out_2.bias", {n2}, &b2)) return 1;
        m->dense3.clear();
        for (int sd = 0; sd < ND; ++sd)
            for (int i = 0; i < nd[sd]; ++i) {
                int row0, nrows;
                split_range(dst[sd].rows, nd[sd], i, &row0, &nrows);
                const int kdim = dst[sd].kdim, kp = round_up(kdim, 4), RP = round_up(nrows, 8);
                const float* src = sd == 0 ? skw->data.data() : (sd == 1 ? w1->data.data() : w2->data.data());
                swn_model::D3Cta dc;
                dc.w_off = alloc((size_t)kp * RP);
                for (int r = 0; r < nrows; ++r)
                    for (int k = 0; k < kdim; ++k) W[dc.w_off + (size_t)k * RP + r] = src[(size_t)(row0 + r) * kdim + k];
                dc.b_off = alloc(nrows);
                for (int r = 0; r < nrows; ++r)
                    W[dc.b_off + r] = sd == 0 ? W[m->cs_off3 + row0 + r] : (sd == 1 ? b1->data[row0 + r] : b2->data[row0 + r]);
                m->dense3.push_back(dc);
            }
        if (upload(m->wbuf3, W)) return 1;
        m->dense3_key = dkey;
        m->plan3_key = -1;
    }
    {
        int cta = L * geff, di = 0;
        for (int sd = 0; sd < ND; ++sd)
            for (int i = 0; i < nd[sd]; ++i, ++cta, ++di) { desc[cta].w_off = m->dense3[di].w_off; desc[cta].b_off = m->dense3[di].b_off; }
    }
    const int NC = BS * seg;
    const long long key = ((long long)B << 32) | ((long long)BS << 16) | ((long long)geff << 8) | ((long long)NT << 2) | ((long long)(tinst_sel - 1) << 1) | (umma ? 1 : 0);
    if (key != m->plan3_key) {
        CUDA_OK(cudaStreamSynchronize(st));
        if (upload(m->desc3, desc)) return 1;
        m->plan3_key = key;
    }
    const size_t n_cnt = (size_t)(L + ND + 1) * M * swn::B3_CNT_STRIDE;
    const size_t n_skip = (size_t)std::max(L - 1, 1) * M * S * BS;
    if (m->tiles3.ensure(tiles_total) || m->cnt3.ensure(n_cnt) || m->skip3.ensure(n_skip) ||
        m->hist3.ensure((size_t)M * BS * swn::B3_HIST) ||
        m->ring3.ensure(std::max<size_t>(ring_total, 4))) return 1;
    // loop steps every slot has to run: the longest of its utterances (ragged batches leave the ring early)
    m->slot3_h.assign(M, (int32_t)n_iter);
    if (a->n_samples_list) {
        for (int s2 = 0; s2 < M; ++s2) {
            int64_t need = 1;
            for (int b = s2 * BS; b < std::min(B, (s2 + 1) * BS); ++b)
                need = std::max<int64_t>(need, (a->n_samples_list[b] + seg - 1) / seg);
            m->slot3_h[s2] = (int32_t)std::min<int64_t>(need, n_iter);
        }
    }
    if (m->slot3.ensure(M)) return 1;
    CUDA_OK(cudaStreamSynchronize(st));                  // the previous launch may still read the old steps
    CUDA_OK(cudaMemcpyAsync(m->slot3.p, m->slot3_h.data(), M * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(m->cnt3.p, 0, n_cnt * sizeof(unsigned int), st));
    if (umma) CUDA_OK(cudaMemsetAsync(m->tiles3.p, 0, tiles_total * sizeof(float), st));     // padding columns of the tiles must be finite
    if (ring_total) CUDA_OK(cudaMemsetAsync(m->ring3.p, 0, ring_total * sizeof(float), st));
    CUDA_OK(cudaMemsetAsync(m->status.p, 0, 8 * sizeof(int32_t), st));

    swn::BatchParams P;
    memset(&P, 0, sizeof(P));
    P.kind = m->cfg.kind; P.H = H; P.S = S; P.K = K; P.L = L; P.lpc = m->lpc;
    P.n1 = n1; P.n1p = m->n1p; P.n2 = n2; P.n2p = n2p; P.Q = m->Q; P.U = m->U; P.R = m->R; P.B = B; P.M = M; P.Tf = Tf;
    const int rem = (m->R + 1) % seg;
    P.t_first = rem == 0 ? 0 : rem - seg;
    P.it0 = (m->R + 1 - P.t_first) / seg - 1;
    P.n_iter = (int)n_iter;
    P.n_it = P.it0 + P.n_iter;
    P.p1_split = M >= 16 ? 1 : 0;
    if (getenv("SWN_P1_SPLIT")) P.p1_split = atoi(getenv("SWN_P1_SPLIT")) ? 1 : 0;
    P.NT = NT; P.tail_inst = tinst_sel; P.tail_gemm = tail_gemm ? 1 : 0;
    P.tail_cnt = L + ND; P.tail_nprod = nd[ND - 1];
    P.tail_in_off = tile_off[L + ND]; P.tail_slot_f = dst[ND - 1].rows * BS;
    P.n_out = n_iter * seg;
    P.desc = m->desc3.p; P.wbuf = m->wbuf3.p;
    P.P = m->ptab.p; P.pconst = m->pconst.p; P.gain = m->gain.p;
    P.in_tab = m->in_tab.p; P.in_bias = m->in_bias.p;
    P.w2_off = m->w2_off3; P.b2_off = m->b2_off3; P.cs_off = m->cs_off3;
    // cell 0 building its own input tile (builder warps / the drain warps in tensor-core mode) measured SLOWER than the
    // tail doing it (few warps, serial dependent code: ~9 us per tile against ~4 us): only when SWN_BUILD_X=1
    P.hist_g = m->hist3.p; P.build_x = (build_x && build_env) ? 1 : 0;
    P.umma = umma ? 1 : 0; P.NP = NPsel; P.ndb = ndb_sel;
    P.slot_nit = m->slot3.p;
    P.tiles = m->tiles3.p; P.cnt = m->cnt3.p; P.skipbuf = m->skip3.p; P.ring_scratch = m->ring3.p;
    P.noise = a->noise_dev; P.seed = a->seed; P.forced = a->forced_dev; P.seedw = a->seed_dev; P.out = a->out_dev; P.head_out = a->head_dev;
    P.status = m->status.p;
    P.timeout_cycles = 4000000000ULL;
    P.prof = (unsigned long long*)a->prof_dev; P.prof_start = a->prof_start; P.prof_count = a->prof_dev ? a->prof_count : 0;
    (void)NC;

    CUDA_OK(cudaEventRecord(m->ev[2], st));
    cudaError_t e = fn(P, grid, smem_max, st);
    if (e != cudaSuccess) return fail("batch kernel launch failed: %s (grid %d, %zu B smem)", cudaGetErrorString(e), grid, smem_max);
    CUDA_OK(cudaEventRecord(m->ev[3], st));
    m->gen_timed = true;
    m->info.kernel_launches++;
    m->info.smem_bytes = (int32_t)smem_max;
    m->info.batch_tile = BS;
    m->info.n_slots = M;
    m->info.grid = grid;
    m->info.kernel_variant = umma ? 3 : 4;          // 3: tensor-core cells (tcgen05), 4: fp32 FFMA cells
    m->info.ctas_per_layer = geff;
    m->info.head_ctas = grid - L * geff;
    return 0;
}

// ---- packing of one forward GEMM operand (weights W[rows][ntaps][C] given through `at(row, tap, c)`)
template <typename F>
static int pack_fwd(swn_model::FwdW& w, int rows, int ntaps, int C, bool gate, int H, const std::vector<float>& bias, F at) {
    const int BK = swn::fwd_bk(), BM = swn::fwd_bm();
    w.rows = rows; w.ntaps = ntaps; w.C = C; w.Cpad = round_up(C, BK);
    const int ktot = ntaps * w.Cpad;
    w.n_kb = ktot / BK;
    w.n_mt = gate ? (H + BM / 2 - 1) / (BM / 2) : (rows + BM - 1) / BM;
    std::vector<float> W((size_t)rows * ktot, 0.f);
    for (int r = 0; r < rows; ++r)
        for (int j = 0; j < ntaps; ++j)
            for (int c = 0; c < C; ++c) W[(size_t)r * ktot + (size_t)j * w.Cpad + c] = at(r, j, c);
    std::vector<int> map((size_t)w.n_mt * BM, -1);
    for (int i = 0; i < w.n_mt * BM; ++i) {
        if (gate) { const int pair = (i / BM) * (BM / 2) + (i % BM) / 2; if (pair < H) map[i] = (i & 1) * H + pair; }
        else if (i < rows) map[i] = i;
    }
    std::vector<float> packed((size_t)w.n_mt * w.n_kb * 2 * BM * BK);
    swn::pack_fwd_weights(W.data(), ktot, w.n_mt, map.data(), packed.data());
    if (upload(w.A, packed) || upload(w.bias, bias)) return 1;
    return 0;
}

static int fwd_prepare(swn_model* m) {
    if (m->fwd_ready) return 0;
    const int H = m->H, S = m->S, K = m->K, L = m->L, n1 = m->n1, n2 = m->n2;
    m->fwd_gate.clear();
    m->fwd_gate.resize(L);
    const HostTensor *x, *y;
    for (int l = 0; l < L; ++l) {
        if (need(m, "dil_h." + std::to_string(l) + ".conv.weight", {2 * H, H, K}, &x) ||
            need(m, "dil_h." + std::to_string(l) + ".conv.bias", {2 * H}, &y)) return 1;
        auto& w = m->fwd_gate[l];
        if (pack_fwd(w, 2 * H, K, H, true, H, y->data, [&](int r, int j, int c) { return x->data[((size_t)r * H + c) * K + j]; })) return 1;
        w.delay.resize(K);
        for (int j = 0; j < K; ++j) w.delay[j] = (K - 1 - j) * m->dil[l];
    }
    {   // sum of the L skip projections as one contraction over the stacked layer outputs (cswnv_shift1.py:209-225)
        std::vector<const HostTensor*> sw(L);
        std::vector<float> bias(S, 0.f);
        for (int l = 0; l < L; ++l) {
            if (need(m, "out_skip." + std::to_string(l) + ".weight", {S, H, 1}, &sw[l]) ||
                need(m, "out_skip." + std::to_string(l) + ".bias", {S}, &y)) return 1;
            for (int r = 0; r < S; ++r) bias[r] += y->data[r];
        }
        if (pack_fwd(m->fwd_skip, S, L, H, false, H, bias, [&](int r, int j, int c) { return sw[j]->data[(size_t)r * H + c]; })) return 1;
        m->fwd_skip.delay.assign(L, 0);
    }
    if (need(m, "out_1.weight", {n1, S, 1}, &x) || need(m, "out_1.bias", {n1}, &y)) return 1;
    if (pack_fwd(m->fwd_out1, n1, 1, S, false, H, y->data, [&](int r, int, int c) { return x->data[(size_t)r * S + c]; })) return 1;
    m->fwd_out1.delay.assign(1, 0);
    if (need(m, "out_2.weight", {n2, n1, 1}, &x) || need(m, "out_2.bias", {n2}, &y)) return 1;
    if (pack_fwd(m->fwd_out2, n2, 1, n1, false, H, y->data, [&](int r, int, int c) { return x->data[(size_t)r * n1 + c]; })) return 1;
    m->fwd_out2.delay.assign(1, 0);
    m->fwd_ready = true;
    return 0;
}

extern "C" {

const char* swn_last_error(void) { return g_err; }

int swn_create(const swn_config* cfg, int device, swn_handle* out) {
    if (!cfg || !out) return fail("swn_create: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("swn_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail("swn_create: device %d out of range (have %d)", device, ndev);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail("swn_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                    prop.minor);
    if (cfg->kind != SWN_KIND_CSWNV && cfg->kind != SWN_KIND_DSWNV) return fail("swn_create: unknown kind %d", cfg->kind);
    swn_model* m = new swn_model();
    m->cfg = *cfg;
    m->device = device;
    m->n_sm = prop.multiProcessorCount;
    const bool is_c = cfg->kind == SWN_KIND_CSWNV;
    m->H = cfg->hid_chn; m->S = cfg->skip_chn; m->K = cfg->kernel_size;
    m->L = cfg->dilation_depth * cfg->dilation_repeat;
    m->seg = is_c ? cfg->seg : 1;
    m->lpc = is_c ? cfg->lpc : 0;
    m->Q = cfg->n_quantize; m->U = cfg->upsampling_factor;
    m->n1 = is_c ? m->S : m->Q;
    m->n2 = is_c ? 2 * m->seg + m->lpc : m->Q;
    m->Hp = round_up(m->H, 64); m->Sp = round_up(m->S, 64); m->n1p = round_up(m->n1, 64);
    m->NCH = m->Hp / 64;
    m->Ca = cfg->n_aux * ipow(cfg->aux_kernel_size, cfg->aux_dilation_size);
    m->Cap = round_up(m->Ca, 16);
    m->NP = m->L * m->seg * 2 * m->H;
    int R = m->K - 1;
    for (int r = 0; r < cfg->dilation_repeat; ++r)
        for (int d = 0; d < cfg->dilation_depth; ++d) {
            m->dil.push_back(ipow(m->K, d));
            R += ipow(m->K, d + 1) - ipow(m->K, d);
        }
    m->R = R;
    const char* bad = nullptr;
    if (m->H < 4 || m->S < 4 || m->K < 2 || m->L < 1 || m->U < 1 || cfg->n_aux < 1) bad = "non-positive dimension";
    else if (m->NCH > 4) bad = "hid_chn > 256 is not instantiated";
    else if (m->seg < 1 || m->K + m->seg > 32) bad = "kernel_size + seg > 32";
    else if (m->lpc < 0 || m->lpc > 8) bad = "lpc > 8";
    else if (!is_c && m->Q != 256) bad = "n_quantize must be 256 (the reference's OneHot is fixed at depth 256, dswnv.py:68)";
    else if (is_c && cfg->audio_in_flag) bad = "audio_in_flag is a DSWNV option";
    else if (swn::batch_tiles_for_seg(m->seg)[0] == 0) bad = "seg not in {1,2,3,4,5,10}";
    else if (cfg->aux_kernel_size % 2 == 0) bad = "aux_kernel_size must be odd";
    if (bad) { int rc = fail("swn_create: unsupported configuration: %s", bad); delete m; return rc; }
    m->NH = 4;
    m->G = (m->n_sm - m->NH) / m->L;
    if (m->G < 1) { int rc = fail("swn_create: %d SMs cannot host %d layers", m->n_sm, m->L); delete m; return rc; }
    m->Geff = std::min(m->G, m->H);
    if ((m->H + m->Geff - 1) / m->Geff > 2 * SWN_NWARPS) {
        int rc = fail("swn_create: %d gate pairs per CTA exceed %d", (m->H + m->Geff - 1) / m->Geff, 2 * SWN_NWARPS);
        delete m; return rc;
    }
    memset(&m->info, 0, sizeof(m->info));
    m->info.abi_version = SWN_ABI_VERSION;
    m->info.receptive_field = m->R;
    m->info.eff_seg = m->seg;
    m->info.head_dim = m->n2;
    m->info.n_sm = m->n_sm;
    m->info.ctas_per_layer = m->Geff;
    m->info.head_ctas = m->NH;
    *out = m;
    return 0;
}

int swn_set_tensor(swn_handle h, const char* key, const float* host, const int64_t* shape, int32_t ndim) {
    if (!h || !key || !host || (!shape && ndim > 0)) return fail("swn_set_tensor: null argument");
    if (ndim < 0 || ndim > 8) return fail("swn_set_tensor('%s'): ndim %d out of range", key, ndim);
    HostTensor x;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] <= 0 || shape[i] > (int64_t)1 << 31 || n * (size_t)shape[i] > ((size_t)1 << 32))
            return fail("swn_set_tensor('%s'): dimension %d has extent %lld", key, i, (long long)shape[i]);
        x.shape.push_back(shape[i]); n *= (size_t)shape[i];
    }
    x.data.assign(host, host + n);
    h->t[key] = std::move(x);
    h->finalized = false;
    h->fwd_ready = false;
    return 0;
}

int swn_finalize(swn_handle m) {
    if (!m) return fail("swn_finalize: null handle");
    CUDA_OK(cudaSetDevice(m->device));
    const swn_config& c = m->cfg;
    const bool is_c = c.kind == SWN_KIND_CSWNV;
    const int H = m->H, Hp = m->Hp, S = m->S, Sp = m->Sp, K = m->K, L = m->L, seg = m->seg, Q = m->Q, U = m->U;
    const int Ca = m->Ca, n1 = m->n1, n1p = m->n1p, n2 = m->n2, twoH = 2 * H;
    const HostTensor *x, *y;

    // ---------------- conditioning weights: scale_in + conv_aux chain (cswnv_shift1.py:152-154)
    for (auto& cw : m->aux_convs) { cw.w_t.release(); cw.b.release(); }
    m->aux_convs.clear();
    m->aux_convs.resize(1 + c.aux_dilation_size);
    {
        if (need(m, "scale_in.weight", {c.n_aux, c.n_aux, 1}, &x) || need(m, "scale_in.bias", {c.n_aux}, &y)) return 1;
        auto& cw = m->aux_convs[0];
        cw.Ci = c.n_aux; cw.Co = c.n_aux; cw.ks = 1; cw.dil = 1; cw.pad = 0;
        std::vector<float> wt((size_t)c.n_aux * c.n_aux);
        for (int co = 0; co < c.n_aux; ++co)
            for (int ci = 0; ci < c.n_aux; ++ci) wt[(size_t)ci * c.n_aux + co] = x->data[(size_t)co * c.n_aux + ci];
        if (upload(cw.w_t, wt) || upload(cw.b, y->data)) return 1;
    }
    for (int i = 0; i < c.aux_dilation_size; ++i) {
        const int ks = c.aux_kernel_size;
        const int Ci = c.n_aux * ipow(ks, i), Co = c.n_aux * ipow(ks, i + 1);
        const std::string base = "conv_aux.conv." + std::to_string(i);
        if (need(m, base + ".weight", {Co, Ci, ks}, &x) || need(m, base + ".bias", {Co}, &y)) return 1;
        auto& cw = m->aux_convs[i + 1];
        cw.Ci = Ci; cw.Co = Co; cw.ks = ks; cw.dil = ipow(ks, i); cw.pad = (ipow(ks, i + 1) - ipow(ks, i)) / 2;
        std::vector<float> wt((size_t)Ci * ks * Co);
        for (int co = 0; co < Co; ++co)
            for (int ci = 0; ci < Ci; ++ci)
                for (int kk = 0; kk < ks; ++kk)
                    wt[((size_t)ci * ks + kk) * Co + co] = x->data[((size_t)co * Ci + ci) * ks + kk];
        if (upload(cw.w_t, wt) || upload(cw.b, y->data)) return 1;
    }
    // upsampling: per-phase gain + bias (cswnv_shift1.py:48-51)
    const HostTensor *up_w, *up_b;
    if (need(m, "upsampling.conv.weight", {1, 1, 1, U}, &up_w) || need(m, "upsampling.conv.bias", {1}, &up_b)) return 1;
    if (upload(m->gain, up_w->data)) return 1;
    const double b_u = up_b->data[0];

    // in_x of every layer -> projection matrix Wp[Cap][NP] and constants (DESIGN.md "conditioning hoist")
    const bool conv2d = is_c && c.aux_conv2d_flag && seg > 1;
    const int A_in = conv2d ? Ca : Ca * seg;                    // in_x in-channels from the aux path
    const int A_tot = A_in + ((!is_c && c.audio_in_flag) ? Q : 0);
    const HostTensor *w2d = nullptr, *b2d = nullptr;
    if (conv2d && (need(m, "aux_conv2d.weight", {Ca, Ca, seg, 1}, &w2d) || need(m, "aux_conv2d.bias", {Ca}, &b2d))) return 1;
    std::vector<float> wp((size_t)m->Cap * m->NP, 0.f), pconst((size_t)L * twoH), xa;
    if (!is_c && c.audio_in_flag) xa.assign((size_t)L * Q * twoH, 0.f);
    std::vector<double> weff((size_t)Ca * seg);
    for (int l = 0; l < L; ++l) {
        const std::string base = "in_x." + std::to_string(l);
        if (need(m, base + ".weight", {twoH, A_tot, 1}, &x) || need(m, base + ".bias", {twoH}, &y)) return 1;
        for (int row = 0; row < twoH; ++row) {
            const float* wr = x->data.data() + (size_t)row * A_tot;
            double beff = y->data[row];
            if (conv2d) {
                // in_x(aux_conv2d(unfold(x)))  (cswnv_shift1.py:306-308): compose the two affine maps
                std::fill(weff.begin(), weff.end(), 0.0);
                for (int c2 = 0; c2 < Ca; ++c2) {
                    const double w = wr[c2];
                    beff += w * b2d->data[c2];
                    const float* src = w2d->data.data() + (size_t)c2 * Ca * seg;   // [c][s]
                    for (int i = 0; i < Ca * seg; ++i) weff[i] += w * src[i];
                }
            } else {
                for (int i = 0; i < Ca * seg; ++i) weff[i] = wr[i];                // channel index c*seg+s (cswnv:310-311)
            }
            double wsum = 0.0;
            for (int cc = 0; cc < Ca; ++cc)
                for (int s = 0; s < seg; ++s) {
                    const double w = weff[(size_t)cc * seg + s];
                    wsum += w;
                    wp[(size_t)cc * m->NP + ((size_t)l * seg + s) * twoH + row] = (float)w;
                }
            pconst[(size_t)l * twoH + row] = (float)(beff + b_u * wsum);
            if (!xa.empty())
                for (int q = 0; q < Q; ++q) xa[((size_t)l * Q + q) * twoH + row] = wr[A_in + q];
        }
    }
    if (upload(m->wp, wp) || upload(m->pconst, pconst)) return 1;
    if (!xa.empty() && upload(m->xa_tab, xa)) return 1;

    // ---------------- input stage tables (causal conv over the audio, cswnv_shift1.py:164-167,352; dswnv.py:222-226,345)
    {
        const HostTensor *wc, *bc, *ww = nullptr, *bw = nullptr;
        const int cin = c.wav_conv_flag ? H : (is_c ? 1 : Q);
        if (need(m, "causal.conv.weight", {H, cin, K}, &wc) || need(m, "causal.conv.bias", {H}, &bc)) return 1;
        if (c.wav_conv_flag && (need(m, "wav_conv.weight", {H, is_c ? 1 : Q, 1}, &ww) || need(m, "wav_conv.bias", {H}, &bw)))
            return 1;
        std::vector<float> tab;
        if (is_c) {
            tab.assign((size_t)K * 2 * H, 0.f);
            for (int j = 0; j < K; ++j)
                for (int r = 0; r < H; ++r) {
                    double cb = 0.0, v = 0.0;
                    if (c.wav_conv_flag) {
                        for (int k = 0; k < H; ++k) {
                            const double w = wc->data[((size_t)r * H + k) * K + j];
                            v += w * ww->data[k]; cb += w * bw->data[k];
                        }
                    } else {
                        v = wc->data[(size_t)r * K + j];
                    }
                    tab[((size_t)j * 2) * H + r] = (float)cb;
                    tab[((size_t)j * 2 + 1) * H + r] = (float)v;
                }
        } else {
            tab.assign((size_t)K * Q * H, 0.f);
            for (int j = 0; j < K; ++j)
                for (int r = 0; r < H; ++r) {
                    if (c.wav_conv_flag) {
                        double cb = 0.0;
                        for (int k = 0; k < H; ++k) cb += (double)wc->data[((size_t)r * H + k) * K + j] * bw->data[k];
                        for (int q = 0; q < Q; ++q) {
                            double v = cb;
                            for (int k = 0; k < H; ++k)
                                v += (double)wc->data[((size_t)r * H + k) * K + j] * ww->data[(size_t)k * Q + q];
                            tab[((size_t)j * Q + q) * H + r] = (float)v;
                        }
                    } else {
                        for (int q = 0; q < Q; ++q) tab[((size_t)j * Q + q) * H + r] = wc->data[((size_t)r * Q + q) * K + j];
                    }
                }
        }
        if (upload(m->in_tab, tab) || upload(m->in_bias, bc->data)) return 1;
        m->in_tab_h = tab;
    }

    // ---------------- in-loop weights of the layer-group kernel (v1), packed per CTA
    if (pack_v1(m, std::min(m->G, m->H))) return 1;
    std::vector<const HostTensor*> skw(L), skb(L);
    for (int l = 0; l < L; ++l) {
        if (need(m, "out_skip." + std::to_string(l) + ".weight", {S, H, 1}, &skw[l]) ||
            need(m, "out_skip." + std::to_string(l) + ".bias", {S}, &skb[l])) return 1;
    }
    // ---------------- v2 packing: the 16 CTAs of the core cluster (swn_core.cuh)
    m->core_ok = false;
    m->core_plan_key = -1;
    m->g3 = 0; m->plan3_key = -1;
    {
        const bool audio_in = (!is_c && c.audio_in_flag);
        int nct = 0;
        m->tap_base.assign(L, 0);
        m->ncore.assign(L, 1);
        m->hl.assign(L, swn_model::HelpLayer());
        bool help_lanes_ok = true;
        for (int l = 0; l < L; ++l) {
            m->tap_base[l] = nct;
            const int d = m->dil[l];
            // one sample per step: delay 0, and delay 1 of the dilation-1 layers (two steps of helper slack);
            // seg > 1 (swn_coreseg.cuh): every tap whose delay is shorter than a step (one step of helper slack)
            const int ncore = seg > 1 ? std::min(K, (seg + d - 1) / d) : ((d == 1 && K >= 2) ? 2 : 1);
            m->ncore[l] = ncore;
            nct += ncore;
            if (seg > 1 && (K - ncore) * seg > 32) help_lanes_ok = false;
            auto& hl = m->hl[l];
            for (int j = K - 1 - ncore; j >= 0; --j) {          // increasing delay
                hl.tapj.push_back(j);
                hl.delay.push_back((K - 1 - j) * d);
            }
            hl.ntaps = (int)hl.tapj.size();
            hl.dmin = hl.ntaps ? hl.delay[0] : 0;
        }
        m->n_core_taps = nct;
        const int tcols = 2 * 2 * m->NCH;
        const bool fits1 = seg == 1 && !audio_in && H <= 256 && S <= 256 && n1 <= 256 && K <= swn::CORE_KMAX && L <= 16 &&
                           nct * tcols <= 128 && (is_c ? n2 <= 16 : n2 <= 256);
        // seg > 1: 4 NCH^2 tensor-memory columns per tap and thread, one (tap, column) per helper lane
        const bool fits5 = is_c && seg >= 2 && seg <= 5 && m->NCH <= 3 && H <= 192 && S <= 256 && n1 <= 256 && K <= swn::CORE_KMAX &&
                           L <= 16 && nct * 4 * m->NCH * m->NCH <= 512 && n2 <= 16 && help_lanes_ok;
        const bool fits = fits1 || fits5;
        if (fits) {
            std::vector<float>& V = m->wbuf2_h;
            V.clear();
            auto alloc2 = [&](size_t n) { size_t off = (V.size() + 3) & ~size_t(3); V.resize(off + n, 0.f); return (int64_t)off; };
            m->ccta.assign(swn::CORE_CL, swn::CoreCta());
            std::vector<const HostTensor*> dw(L), db(L);
            for (int l = 0; l < L; ++l)
                if (need(m, "dil_h." + std::to_string(l) + ".conv.weight", {twoH, H, K}, &dw[l]) ||
                    need(m, "dil_h." + std::to_string(l) + ".conv.bias", {twoH}, &db[l])) return 1;
            const HostTensor *w1, *b1, *w2, *b2;
            if (need(m, "out_1.weight", {n1, S, 1}, &w1) || need(m, "out_1.bias", {n1}, &b1) ||
                need(m, "out_2.weight", {n2, n1, 1}, &w2) || need(m, "out_2.bias", {n2}, &b2)) return 1;
            for (int r = 0; r < swn::CORE_CL; ++r) {
                swn::CoreCta& q = m->ccta[r];
                memset(&q, 0, sizeof(q));
                split_range(H, swn::CORE_CL, r, &q.pair0, &q.npairs);
                split_range(S, swn::CORE_CL, r, &q.skip0, &q.nskip);
                split_range(n1, swn::CORE_CL, r, &q.b0, &q.nb);
                if (is_c) { q.c0 = 0; q.nc = n2; } else split_range(n2, swn::CORE_CL, r, &q.c0, &q.nc);
                q.i0 = q.pair0; q.ni = q.npairs;
                const int np = q.npairs;
                q.tap_off = alloc2((size_t)std::max(nct * np, 1) * 2 * Hp);
                q.bias_off = alloc2((size_t)std::max(L * np, 1) * 2);
                for (int l = 0; l < L; ++l) {
                    const int ncore = (l + 1 < L ? m->tap_base[l + 1] : nct) - m->tap_base[l];
                    for (int e = 0; e < ncore; ++e) {
                        const int j = K - 1 - e, ct = m->tap_base[l] + e;
                        for (int pi = 0; pi < np; ++pi)
                            for (int zc = 0; zc < 2; ++zc) {
                                const int row = zc * H + q.pair0 + pi;
                                float* dst = V.data() + q.tap_off + ((size_t)(ct * np + pi) * 2 + zc) * Hp;
                                for (int k = 0; k < H; ++k) dst[k] = dw[l]->data[((size_t)row * H + k) * K + j];
                            }
                    }
                    for (int pi = 0; pi < np; ++pi) {
                        V[q.bias_off + (size_t)(l * np + pi) * 2] = db[l]->data[q.pair0 + pi];
                        V[q.bias_off + (size_t)(l * np + pi) * 2 + 1] = db[l]->data[H + q.pair0 + pi];
                    }
                }
                q.skip_off = alloc2((size_t)std::max(L * q.nskip, 1) * Hp);
                q.skipc_off = alloc2((size_t)std::max(q.nskip, 1));
                for (int l = 0; l < L; ++l)
                    for (int rr = 0; rr < q.nskip; ++rr)
                        for (int k = 0; k < H; ++k)
                            V[q.skip_off + ((size_t)l * q.nskip + rr) * Hp + k] = skw[l]->data[(size_t)(q.skip0 + rr) * H + k];
                for (int rr = 0; rr < q.nskip; ++rr) {
                    double cs = 0.0;
                    for (int l = 0; l < L; ++l) cs += skb[l]->data[q.skip0 + rr];
                    V[q.skipc_off + rr] = (float)cs;
                }
                q.w1_off = alloc2((size_t)std::max(q.nb, 1) * Sp);
                q.b1_off = alloc2((size_t)std::max(q.nb, 1));
                for (int rr = 0; rr < q.nb; ++rr) {
                    for (int k = 0; k < S; ++k) V[q.w1_off + (size_t)rr * Sp + k] = w1->data[(size_t)(q.b0 + rr) * S + k];
                    V[q.b1_off + rr] = b1->data[q.b0 + rr];
                }
                q.w2_off = alloc2((size_t)std::max(q.nc, 1) * n1p);
                q.b2_off = alloc2((size_t)std::max(q.nc, 1));
                for (int rr = 0; rr < q.nc; ++rr) {
                    for (int k = 0; k < n1; ++k) V[q.w2_off + (size_t)rr * n1p + k] = w2->data[(size_t)(q.c0 + rr) * n1 + k];
                    V[q.b2_off + rr] = b2->data[q.c0 + rr];
                }
                const int nip = (q.ni + 3) & ~3;
                q.tin_off = alloc2((size_t)std::max(Q, 1) * std::max(nip, 4));
                if (!is_c)
                    for (int qq = 0; qq < Q; ++qq)
                        for (int rr = 0; rr < q.ni; ++rr)
                            V[q.tin_off + (size_t)qq * nip + rr] = m->in_tab_h[((size_t)(K - 1) * Q + qq) * H + q.i0 + rr];
            }
            m->wbuf2_core_floats = (V.size() + 3) & ~size_t(3);
            V.resize(m->wbuf2_core_floats, 0.f);
            m->core_ok = true;
        }
    }
    if (m->status.ensure(8)) return 1;
    for (int i = 0; i < 4; ++i)
        if (!m->ev[i]) CUDA_OK(cudaEventCreate(&m->ev[i]));
    m->finalized = true;
    m->condB = 0;
    m->plan_key = -1;
    return 0;
}

int swn_condition(swn_handle m, const float* aux_dev, int32_t B, int32_t T_f, void* stream) {
    if (!m || !aux_dev) return fail("swn_condition: null argument");
    if (!m->finalized) return fail("swn_condition: call swn_finalize first");
    if (B < 1 || T_f < 1) return fail("swn_condition: B=%d T_f=%d", B, T_f);
    CUDA_OK(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    size_t amax = 0;
    for (auto& cw : m->aux_convs) amax = std::max(amax, (size_t)std::max(cw.Co, m->Cap));
    if (m->act0.ensure((size_t)B * amax * T_f) || m->act1.ensure((size_t)B * amax * T_f)) return 1;
    if (m->ptab.ensure((size_t)B * T_f * m->NP)) return 1;
    CUDA_OK(cudaEventRecord(m->ev[0], st));
    const float* cur = aux_dev;
    float* bufs[2] = {m->act0.p, m->act1.p};
    int which = 0;
    for (size_t i = 0; i < m->aux_convs.size(); ++i) {
        auto& cw = m->aux_convs[i];
        const int last = (i + 1 == m->aux_convs.size());
        CUDA_OK(swn::launch_aux_conv(cur, cw.w_t.p, cw.b.p, bufs[which], B, cw.Ci, cw.Co, T_f, cw.ks, cw.dil, cw.pad, last, st));
        cur = bufs[which];
        which ^= 1;
        m->info.kernel_launches++;
    }
    // cur is now frame-major y^T [B*T_f][Cap]
    CUDA_OK(swn::launch_proj_gemm(cur, m->wp.p, m->ptab.p, (long long)B * T_f, m->NP, m->Cap, st));
    m->info.kernel_launches++;
    CUDA_OK(cudaEventRecord(m->ev[1], st));
    m->cond_timed = true;
    m->condB = B; m->condTf = T_f;
    m->info.cond_bytes = (int64_t)B * T_f * m->NP * 4;
    return 0;
}

int swn_generate(swn_handle m, const swn_gen_args* a, void* stream) {
    if (!m || !a) return fail("swn_generate: null argument");
    if (!m->finalized || m->condB < 1) return fail("swn_generate: call swn_finalize and swn_condition first");
    if (!a->out_dev) return fail("swn_generate: out_dev is null");
    CUDA_OK(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int B = m->condB, Tf = m->condTf, seg = m->seg, H = m->H, Hp = m->Hp, K = m->K, L = m->L, S = m->S;
    const int64_t n_iter = a->n_samples / seg;
    if (n_iter < 1) return fail("swn_generate: n_samples=%lld yields no loop step (seg=%d)", (long long)a->n_samples, seg);
    if (n_iter * seg > (int64_t)Tf * m->U)
        return fail("swn_generate: n_samples=%lld exceeds T_f*upsampling_factor=%lld", (long long)a->n_samples,
                    (long long)Tf * m->U);
    if (n_iter + m->R + 2 > 0x7fffffffLL) return fail("swn_generate: too many steps");

    // ---- small batches, one sample per step: the cluster kernel (v2); everything else: layer groups (v1)
    // 3-7 utterances of a one-sample-per-step model: two at a time through the cluster kernel, one launch after the
    // other (measured, DSWNV H=256: 145 k samples/s against 69-92 k for the layer-group kernel at B=3-4 and 111 k for
    // the batch kernel at B=6; from 8 utterances on the batch kernel wins).  CSWNV seg=1 (the recipe's default model,
    // H=192; tools/gpu_seg1_probe.py, 0.2 s): 24.8 ms per launch of two against 161-165 ms for the batch kernel at any
    // B <= 12 (its ring is latency-bound: 56 us per sample step) - up to 12 utterances the launches of two win.
    // Two utterances: one core cluster EACH where the helpers hold two rings (a cluster steps one column in 6.7 us and two columns
    // in 8.6 us; CSWNV seg 1, H=192: 24.9 -> 19.6 ms).  Three clusters of one utterance are helper-bound on the 64 helper CTAs they
    // leave (28.0 ms against 25.9 ms for two clusters of two), so three utterances take the path below.
    if (m->core_ok && seg == 1 && a->kernel_sel == 0 && a->batch_tile == 0 && !a->prof_dev && B == 2 &&
        m->core_multi1_ok[B - 2] != 0 && !(getenv("SWN_CORE_NCL") && atoi(getenv("SWN_CORE_NCL")) == 1)) {
        const int rc = generate_core(m, a, st, B, Tf, n_iter, 0, true, true, 1, B);
        m->core_multi1_ok[B - 2] = rc == 2 ? 0 : 1;
        if (rc == 0) return 0;
        if (rc == 1) return 1;
        CUDA_OK(cudaStreamSynchronize(st));
    }
    if (m->core_ok && seg == 1 && a->kernel_sel == 0 && a->batch_tile == 0 && !a->prof_dev && B >= 3 &&
        B <= (m->cfg.kind == SWN_KIND_DSWNV ? 7 : 16)) {
        // Two core clusters of two utterances side by side where the helpers' shared memory holds the rings of four utterances
        // (hid_chn <= 192: the CSWNV recipe model; the DSWNV recipe model's helper taps alone need 84 CTAs): launches of
        // up to four, of balanced sizes.
        static const int core_ncl_env = getenv("SWN_CORE_NCL") ? atoi(getenv("SWN_CORE_NCL")) : 0;   // measurements: 1 = launches of two
        int rc = 2;
        if (m->core_multi_ok != 0 && core_ncl_env != 1) {
            const int per = 4, nl = (B + per - 1) / per, base = B / nl, extra = B % nl;
            rc = 0;
            for (int g = 0, b0 = 0; g < nl && rc == 0; ++g) {
                const int Bg = base + (g < extra ? 1 : 0);
                rc = generate_core(m, a, st, Bg, Tf, n_iter, b0, g == 0, g + 1 == nl, 2, 2);
                b0 += Bg;
            }
            m->core_multi_ok = rc == 2 ? 0 : 1;
        }
        if (rc == 2) {
            rc = 0;
            for (int b0 = 0; b0 < B && rc == 0; b0 += 2)
                rc = generate_core(m, a, st, std::min(2, B - b0), Tf, n_iter, b0, b0 == 0, b0 + 2 >= B, 2);
        }
        if (rc == 0) return 0;
        if (rc == 1) return 1;
        CUDA_OK(cudaStreamSynchronize(st));                 // not applicable after all: nothing was launched (rc == 2 comes first)
    }
    if (m->core_ok && seg > 1) {
        // seg > 1: one utterance per core cluster; up to three core clusters run side by side in one launch and share the helper
        // CTAs (two: 80 helpers, three: 64 - the rings of partial sums of the long dilations are per utterance), larger batches
        // as launches of balanced sizes one after the other.  Measured for the recipe model (H=192, seg 5, 0.33 s;
        // profiles/r03_crossover_cswnv.txt): 8.0 / 8.0 / 9.1 ms for 1 / 2 / 3 utterances side by side (two on the three-cluster
        // plan: 8.5 ms), against 33.4 ms for the batch kernel at any B <= 10: up to nine utterances the cluster kernel wins
        // (B = 9: 27.3 ms), four run as 2 + 2 on the two-cluster plan.
        static const int seg_ncl_env = getenv("SWN_SEG_NCL") ? atoi(getenv("SWN_SEG_NCL")) : 0;      // measurements: 1 = one utterance per launch
        const int seg_bmax = 9;
        if (a->kernel_sel == 2 || (a->kernel_sel == 0 && a->batch_tile == 0 && B <= seg_bmax)) {
            int ncl_plan = std::min(B, B == 4 ? 2 : (int)swn::CORE_MAXCL);
            if (seg_ncl_env > 0) ncl_plan = std::min(B, std::min(seg_ncl_env, (int)swn::CORE_MAXCL));
            int rc = 2;
            for (; ncl_plan >= 1 && rc == 2; --ncl_plan) {
                // the fewest launches, of balanced sizes (4 utterances: 2 + 2, not 3 + 1)
                const int nl = (B + ncl_plan - 1) / ncl_plan, base = B / nl, extra = B % nl;
                rc = 0;
                for (int g = 0, b0 = 0; g < nl && rc == 0; ++g) {
                    const int ncl = base + (g < extra ? 1 : 0);
                    rc = generate_coreseg(m, a, st, Tf, n_iter, b0, g == 0, g + 1 == nl, ncl, ncl_plan);
                    b0 += ncl;
                }
            }
            if (rc == 0) return 0;
            if (rc == 1) return 1;
            if (a->kernel_sel == 2) return fail("swn_generate: the cluster kernel is not applicable to this configuration");
            CUDA_OK(cudaStreamSynchronize(st));
        }
    } else if (m->core_ok && a->kernel_sel != 1 && a->kernel_sel < 3) {
        const int rc = generate_core(m, a, st, B, Tf, n_iter);
        if (rc == 0) return 0;
        if (rc == 1) return 1;
        if (a->kernel_sel == 2) return fail("swn_generate: the cluster kernel is not applicable to this configuration/batch");
    } else if (a->kernel_sel == 2) {
        return fail("swn_generate: the cluster kernel is not applicable to this configuration");
    }
    // ---- batches: the pipelined batch kernel (v3)
    if (a->kernel_sel == 3 || a->kernel_sel == 4 || (a->kernel_sel == 0 && B > 4)) {
        const int rc = generate_batch(m, a, st, B, Tf, n_iter);
        if (rc == 0) return 0;
        if (rc == 1) return 1;
        if (a->kernel_sel >= 3) return fail("swn_generate: the batch kernel is not applicable to this configuration/batch");
    }
    // ---- batch tile
    int BS = 0;
    const int* tiles = swn::batch_tiles_for_seg(seg);
    if (a->batch_tile > 0) {
        for (const int* t = tiles; *t; ++t) if (*t == a->batch_tile) BS = *t;
        if (!BS) return fail("swn_generate: batch_tile=%d is not instantiated for seg=%d", a->batch_tile, seg);
    } else {
        BS = tiles[0];
        for (const int* t = tiles; *t; ++t) if (*t >= B) BS = *t;      // smallest tile covering B, else the largest
    }
    const int M = (B + BS - 1) / BS;
    const int NC = BS * seg;
    swn::launch_fn fn = nullptr;
    switch (m->NCH) {
        case 1: fn = swn::find_kernel_nch1(seg, BS); break;
        case 2: fn = swn::find_kernel_nch2(seg, BS); break;
        case 3: fn = swn::find_kernel_nch3(seg, BS); break;
        case 4: fn = swn::find_kernel_nch4(seg, BS); break;
    }
    if (!fn) return fail("swn_generate: no kernel for hid_chn=%d seg=%d batch_tile=%d", H, seg, BS);

    // ---- head groups: the head is a serial chain of ~4 L2 hops per slot, so several independent head groups
    // (each NH CTAs, each serving the slots m % NHG == g) keep the layer groups fed when the batch is pipelined
    int NHG = std::max(1, std::min(7, M));
    int geff = 0;
    for (; NHG >= 1; --NHG) {
        const int G = (m->n_sm - m->NH * NHG) / L;
        geff = std::min(G, H);
        if (G >= 1 && (H + geff - 1) / geff <= 2 * SWN_NWARPS) break;
    }
    if (NHG < 1) return fail("swn_generate: cannot place %d layers on %d SMs", L, m->n_sm);
    if (geff != m->Geff) {
        CUDA_OK(cudaStreamSynchronize(st));
        if (pack_v1(m, geff)) return 1;
    }
    // ---- per-CTA plan: shared memory placement
    const int smem_floats = SWN_MAX_SMEM / 4;
    std::vector<CtaDesc> desc(m->n_sm);
    for (auto& d : desc) { memset(&d, 0, sizeof(d)); d.role = ROLE_IDLE; }
    size_t ring_total = 0;
    int smem_used_max = 0;
    int cta = 0;
    for (int l = 0; l < L; ++l)
        for (int g = 0; g < m->Geff; ++g, ++cta) {
            const auto& lc = m->lcta[l][g];
            CtaDesc& d = desc[cta];
            d.role = ROLE_LAYER; d.layer = l; d.dil = m->dil[l];
            d.pair0 = lc.pair0; d.npairs = lc.npairs; d.skip0 = lc.skip0; d.nskip = lc.nskip;
            d.w_off = lc.w_off; d.b_off = lc.b_off; d.ws_off = lc.ws_off;
            d.ring_len = (K - 1) * d.dil + seg;
            int used = 0;
            d.sm_h = used; used += 2 * NC * Hp;
            const int tap_f = lc.npairs * 2 * Hp;
            d.sm_w = used;
            d.smem_taps = std::max(0, std::min(K, tap_f ? (smem_floats - used) / tap_f : K));
            used += d.smem_taps * tap_f;
            d.sm_ws = used;
            d.skip_in_smem = (d.smem_taps == K && used + lc.nskip * Hp <= smem_floats);
            if (d.skip_in_smem) used += lc.nskip * Hp;
            const size_t ring_f = (size_t)M * d.ring_len * lc.npairs * 2 * BS;
            d.sm_ring = used;
            d.ring_in_smem = (d.skip_in_smem && used + ring_f <= (size_t)smem_floats);
            if (d.ring_in_smem) used += (int)ring_f;
            else { d.ring_off = (int64_t)ring_total; ring_total += (ring_f + 3) & ~size_t(3); }
            smem_used_max = std::max(smem_used_max, used);
        }
    const size_t hist_f = (size_t)M * BS * (32 + 8);
    for (int hgi = 0; hgi < NHG; ++hgi)
    for (int hc = 0; hc < m->NH; ++hc, ++cta) {
        const auto& q = m->hcta[hc];
        CtaDesc& d = desc[cta];
        d.role = ROLE_HEAD; d.layer = hc; d.hgroup = hgi;
        d.a0 = q.a0; d.na = q.na; d.b0 = q.b0; d.nb = q.nb; d.c0 = q.c0; d.nc = q.nc; d.i0 = q.i0; d.ni = q.ni;
        d.wa_off = q.wa_off; d.ca_off = q.ca_off; d.wb_off = q.wb_off; d.bb_off = q.bb_off; d.wc_off = q.wc_off; d.bc_off = q.bc_off;
        int used = 0;
        d.sm_x0 = used; used += BS * Hp;
        d.sm_x1 = used; used += BS * m->Sp;
        d.sm_x2 = used; used += BS * m->n1p;
        d.sm_misc = used;
        used += BS * ((q.na + 3) & ~3) + BS * ((q.nc + 3) & ~3) + (int)hist_f + BS * seg + BS + 8;
        used = (used + 3) & ~3;
        if (used > smem_floats)
            return fail("swn_generate: batch of %d utterances needs %d KB of head-CTA state; split the batch", B, used / 256);
        d.sm_wa = used; d.wa_in_smem = used + q.na * Hp <= smem_floats; if (d.wa_in_smem) used += q.na * Hp;
        d.sm_wb = used; d.wb_in_smem = used + q.nb * m->Sp <= smem_floats; if (d.wb_in_smem) used += q.nb * m->Sp;
        d.sm_wc = used; d.wc_in_smem = used + q.nc * m->n1p <= smem_floats; if (d.wc_in_smem) used += q.nc * m->n1p;
        smem_used_max = std::max(smem_used_max, used);
    }
    const int grid = cta;
    const size_t smem_bytes = (size_t)smem_used_max * 4;

    // the plan only changes with (B, batch tile); re-upload it only then, after the stream drained
    const long long plan_key = ((long long)B << 32) | ((long long)BS << 16) | ((long long)NHG << 8) | (long long)seg;
    if (plan_key != m->plan_key) {
        CUDA_OK(cudaStreamSynchronize(st));
        if (upload(m->desc, desc)) return 1;
        m->plan_key = plan_key;
    }
    // ---- exchange buffers
    const size_t n_ring = (size_t)(L + 1) * M * NC * H;
    const size_t n_skip = (size_t)std::max(L - 1, 1) * M * BS * S;
    const size_t n_hx1 = (size_t)M * BS * S, n_hx2 = (size_t)M * BS * m->n1, n_hx3 = (size_t)M * m->NH * BS * 2;
    const size_t n_idx = (size_t)M * BS;
    const size_t n_ll = n_ring + n_skip + n_hx1 + n_hx2 + n_hx3 + n_idx;
    if (m->ll.ensure(n_ll) || m->ring_scratch.ensure(std::max<size_t>(ring_total, 4))) return 1;
    CUDA_OK(cudaMemsetAsync(m->ll.p, 0, n_ll * sizeof(unsigned long long), st));
    if (ring_total) CUDA_OK(cudaMemsetAsync(m->ring_scratch.p, 0, ring_total * sizeof(float), st));
    CUDA_OK(cudaMemsetAsync(m->status.p, 0, 8 * sizeof(int32_t), st));

    GenParams P;
    memset(&P, 0, sizeof(P));
    P.kind = m->cfg.kind; P.H = H; P.Hp = Hp; P.S = S; P.Sp = m->Sp; P.K = K; P.L = L; P.lpc = m->lpc;
    P.n1 = m->n1; P.n1p = m->n1p; P.n2 = m->n2; P.Q = m->Q; P.U = m->U; P.R = m->R; P.B = B; P.M = M; P.Tf = Tf;
    const int rem = (m->R + 1) % seg;
    P.t_first = rem == 0 ? 0 : rem - seg;
    P.it0 = (m->R + 1 - P.t_first) / seg - 1;
    P.n_iter = (int)n_iter;
    P.n_it = P.it0 + P.n_iter;
    P.NH = m->NH; P.NHG = NHG;
    P.audio_in = (m->cfg.kind == SWN_KIND_DSWNV && m->cfg.audio_in_flag) ? 1 : 0;
    P.n_out = n_iter * seg;
    P.desc = m->desc.p;
    P.wbuf = m->wbuf.p;
    P.P = m->ptab.p; P.pconst = m->pconst.p; P.gain = m->gain.p;
    P.in_tab = m->in_tab.p; P.in_bias = m->in_bias.p; P.xa_tab = m->xa_tab.p;
    P.ring_ll = m->ll.p;
    P.skip_ll = P.ring_ll + n_ring;
    P.hx1 = P.skip_ll + n_skip;
    P.hx2 = P.hx1 + n_hx1;
    P.hx3 = P.hx2 + n_hx2;
    P.idx_ll = P.hx3 + n_hx3;
    P.ring_scratch = m->ring_scratch.p;
    P.noise = a->noise_dev; P.seed = a->seed; P.forced = a->forced_dev; P.seedw = a->seed_dev; P.out = a->out_dev; P.head_out = a->head_dev;
    P.status = m->status.p;
    P.timeout_cycles = 4000000000ULL;
    P.prof = (unsigned long long*)a->prof_dev; P.prof_start = a->prof_start; P.prof_count = a->prof_dev ? a->prof_count : 0;
    P.poll_mode = a->poll_mode == 2 ? 1 : 0;      // measured: pure per-word polling is not slower (profiles/)
    P.poll_sleep_ns = 40;
    m->info.grid = grid;

    CUDA_OK(cudaEventRecord(m->ev[2], st));
    cudaError_t e = fn(P, grid, smem_bytes, st);
    if (e != cudaSuccess) return fail("persistent kernel launch failed: %s (grid %d, %zu B smem)", cudaGetErrorString(e), grid, smem_bytes);
    CUDA_OK(cudaEventRecord(m->ev[3], st));
    m->gen_timed = true;
    m->info.kernel_launches++;
    m->info.smem_bytes = (int32_t)smem_bytes;
    m->info.batch_tile = BS;
    m->info.n_slots = M;
    m->info.kernel_variant = 1;
    m->info.ctas_per_layer = m->Geff;
    m->info.head_ctas = m->NH;
    return 0;
}

int swn_teacher_forced(swn_handle m, const void* samples_dev, int64_t n_samples, float* head_dev, void* stream) {
    if (!m || !samples_dev || !head_dev) return fail("swn_teacher_forced: null argument");
    if (!m->finalized) return fail("swn_teacher_forced: call swn_finalize first");
    if (m->condB <= 0) return fail("swn_teacher_forced: call swn_condition first");
    CUDA_OK(cudaSetDevice(m->device));
    const int64_t n_gen = (n_samples / m->seg) * m->seg;
    const size_t esz = m->cfg.kind == SWN_KIND_CSWNV ? sizeof(float) : sizeof(long long);
    if (m->tf_out.ensure((size_t)m->condB * (size_t)std::max<int64_t>(n_gen, 1) * esz)) return 1;   // the sampler's own draws, discarded
    swn_gen_args a;
    memset(&a, 0, sizeof(a));
    a.n_samples = n_samples; a.forced_dev = samples_dev; a.head_dev = head_dev; a.out_dev = m->tf_out.p;
    return swn_generate(m, &a, stream);
}

int swn_forward(swn_handle m, const swn_fwd_args* a, void* stream) {
    if (!m || !a) return fail("swn_forward: null argument");
    if (!m->finalized || m->condB < 1) return fail("swn_forward: call swn_finalize and swn_condition first");
    if (!a->audio_dev || !a->out_dev) return fail("swn_forward: audio_dev / out_dev is null");
    CUDA_OK(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool is_c = m->cfg.kind == SWN_KIND_CSWNV;
    const int B = m->condB, Tf = m->condTf, seg = m->seg, H = m->H, S = m->S, L = m->L, n1 = m->n1, n2 = m->n2, Q = m->Q;
    const long long T_out = (long long)Tf * m->U - 2 * seg + 1;             // cswnv_shift1.py:193-201, dswnv.py:252
    if (T_out < 1) return fail("swn_forward: the conditioning is too short (T_f * U = %lld)", (long long)Tf * m->U);
    if (a->T_a != T_out + seg - 1)
        return fail("swn_forward: audio length %lld does not match the conditioning: expected T_f*U - seg = %lld",
                    (long long)a->T_a, T_out + seg - 1);
    if ((long long)B * T_out > 0x7fffffffLL) return fail("swn_forward: too many time steps in one call");
    if (is_c && (!a->b_dev || !a->log_b_dev)) return fail("swn_forward: b_dev / log_b_dev is null");
    if (L + 1 > 16 || m->K > 16) return fail("swn_forward: more than 16 layers / taps");
    if (!is_c && m->cfg.audio_in_flag && !m->xa_tab.p) return fail("swn_forward: audio_in table missing");
    if (fwd_prepare(m)) return 1;
    const int T = (int)T_out, Ta = (int)a->T_a;
    const size_t rows = (size_t)B * T;
    const int Hq = round_up(H, 4), Sq = round_up(S, 4), n1q = round_up(n1, 4);
    const int ldh = (L + 1) * Hq;
    auto grow = [&](DevBuf<float>& d, size_t n) {
        if (n <= d.n && d.p) return 0;
        if (d.ensure(n)) return 1;
        return cudaMemsetAsync(d.p, 0, n * sizeof(float), st) == cudaSuccess ? 0 : fail("cudaMemsetAsync failed");
    };
    if (grow(m->fwd_h_hi, rows * ldh) || grow(m->fwd_h_lo, rows * ldh) || grow(m->fwd_p, rows * 2 * H) ||
        grow(m->fwd_x1_hi, rows * Sq) || grow(m->fwd_x1_lo, rows * Sq) || grow(m->fwd_x2_hi, rows * n1q) ||
        grow(m->fwd_x2_lo, rows * n1q)) return 1;
    CUDA_OK(cudaMemsetAsync(m->status.p, 0, 8 * sizeof(int32_t), st));
    CUDA_OK(cudaEventRecord(m->ev[2], st));
    const int32_t* idx = nullptr;
    if (!is_c) {
        if (m->fwd_idx.ensure((size_t)B * Ta) || m->fwd_bad.ensure(1)) return 1;
        CUDA_OK(cudaMemsetAsync(m->fwd_bad.p, 0, sizeof(int32_t), st));
        if (a->audio_onehot) CUDA_OK(swn::launch_fwd_onehot((const float*)a->audio_dev, B, Q, Ta, m->fwd_idx.p, m->fwd_bad.p, st));
        else CUDA_OK(swn::launch_fwd_index((const long long*)a->audio_dev, Q, (long long)B * Ta, m->fwd_idx.p, st));
        idx = m->fwd_idx.p;
        m->info.kernel_launches++;
    }
    {
        swn::FwdInput I;
        memset(&I, 0, sizeof(I));
        I.kind = m->cfg.kind; I.H = H; I.K = m->K; I.Q = Q; I.seg = seg; I.T = T; I.Ta = Ta; I.nbatch = B; I.ld = ldh;
        I.in_tab = m->in_tab.p; I.in_bias = m->in_bias.p;
        I.audio_f = is_c ? (const float*)a->audio_dev : nullptr; I.audio_i = idx;
        I.out_hi = m->fwd_h_hi.p; I.out_lo = m->fwd_h_lo.p;
        CUDA_OK(swn::launch_fwd_input(I, st));
        m->info.kernel_launches++;
    }
    auto gemm = [&](const swn_model::FwdW& w, swn::FwdGemm& g) {
        g.A = w.A.p; g.C = w.C; g.Cpad = w.Cpad; g.ntaps = w.ntaps; g.n_kb = w.n_kb; g.n_mt = w.n_mt;
        for (int j = 0; j < w.ntaps; ++j) g.delay[j] = w.delay[j];
        g.T = T; g.nbatch = B; g.rows = w.rows; g.H = H; g.bias = w.bias.p; g.status = m->status.p;
        m->info.kernel_launches++;
        return swn::launch_fwd_gemm(g, st);
    };
    for (int l = 0; l < L; ++l) {
        swn::FwdCond C;
        memset(&C, 0, sizeof(C));
        C.H = H; C.L = L; C.l = l; C.seg = seg; C.U = m->U; C.T = T; C.Tf = Tf; C.Ta = Ta; C.Q = Q; C.nbatch = B;
        C.P = m->ptab.p; C.pconst = m->pconst.p; C.gain = m->gain.p;
        C.xa_tab = (!is_c && m->cfg.audio_in_flag) ? m->xa_tab.p : nullptr; C.audio_i = idx;
        C.out = m->fwd_p.p;
        CUDA_OK(swn::launch_fwd_cond(C, st));
        m->info.kernel_launches++;
        swn::FwdGemm g;
        memset(&g, 0, sizeof(g));
        g.mode = swn::FWD_GATE;
        g.Bhi = m->fwd_h_hi.p; g.Blo = m->fwd_h_lo.p; g.ldb = ldh; g.b_col0 = l * Hq; g.tap_col_stride = 0;
        g.p = m->fwd_p.p; g.carry_col0 = l * Hq;
        g.out_hi = m->fwd_h_hi.p; g.out_lo = m->fwd_h_lo.p; g.ldo = ldh; g.out_col0 = (l + 1) * Hq;
        CUDA_OK(gemm(m->fwd_gate[l], g));
    }
    {
        swn::FwdGemm g;
        memset(&g, 0, sizeof(g));
        g.mode = swn::FWD_RELU;
        g.Bhi = m->fwd_h_hi.p; g.Blo = m->fwd_h_lo.p; g.ldb = ldh; g.b_col0 = Hq; g.tap_col_stride = Hq;
        g.out_hi = m->fwd_x1_hi.p; g.out_lo = m->fwd_x1_lo.p; g.ldo = Sq; g.out_col0 = 0;
        CUDA_OK(gemm(m->fwd_skip, g));
    }
    {
        swn::FwdGemm g;
        memset(&g, 0, sizeof(g));
        g.mode = swn::FWD_RELU;
        g.Bhi = m->fwd_x1_hi.p; g.Blo = m->fwd_x1_lo.p; g.ldb = Sq; g.b_col0 = 0; g.tap_col_stride = 0;
        g.out_hi = m->fwd_x2_hi.p; g.out_lo = m->fwd_x2_lo.p; g.ldo = n1q; g.out_col0 = 0;
        CUDA_OK(gemm(m->fwd_out1, g));
    }
    {
        swn::FwdGemm g;
        memset(&g, 0, sizeof(g));
        g.mode = swn::FWD_LINEAR;
        g.Bhi = m->fwd_x2_hi.p; g.Blo = m->fwd_x2_lo.p; g.ldb = n1q; g.b_col0 = 0; g.tap_col_stride = 0;
        g.out_f = a->out_dev; g.ldf = n2;
        CUDA_OK(gemm(m->fwd_out2, g));
    }
    if (is_c) {
        CUDA_OK(swn::launch_fwd_laplace(a->out_dev, n2, seg, (long long)rows, a->b_dev, a->log_b_dev,
                                        a->clip ? a->b_clip_dev : nullptr, a->clip ? a->log_b_clip_dev : nullptr, st));
        m->info.kernel_launches++;
    }
    CUDA_OK(cudaEventRecord(m->ev[3], st));
    m->gen_timed = true;
    m->info.kernel_variant = 5;
    m->info.grid = (int32_t)(((T + 255) / 256) * B);
    m->info.batch_tile = B; m->info.n_slots = 1;
    return 0;
}

int swn_sync(swn_handle m, void* stream) {
    if (!m) return fail("swn_sync: null handle");
    CUDA_OK(cudaSetDevice(m->device));
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    if (m->cond_timed) { cudaEventElapsedTime(&m->info.last_cond_ms, m->ev[0], m->ev[1]); m->cond_timed = false; }
    if (m->gen_timed) {
        cudaEventElapsedTime(&m->info.last_gen_ms, m->ev[2], m->ev[3]);
        m->gen_timed = false;
        int32_t stv[8];
        CUDA_OK(cudaMemcpy(stv, m->status.p, sizeof(stv), cudaMemcpyDeviceToHost));
        if (stv[0] != 0)
            return fail("%s aborted: watchdog fired at wait site %d (tag %d, CTA %d)", stv[0] == 2 ? "forward GEMM" : "persistent kernel",
                        stv[1], stv[2], stv[3]);
        if (m->info.kernel_variant == 5 && m->fwd_bad.p) {
            int32_t bad = 0;
            CUDA_OK(cudaMemcpy(&bad, m->fwd_bad.p, sizeof(bad), cudaMemcpyDeviceToHost));
            if (bad) return fail("swn_forward: %d columns of the audio tensor are not one-hot", bad);
        }
    }
    return 0;
}

int swn_generate_host(swn_handle m, const float* aux_host, int32_t B, int32_t T_f, int64_t n_samples, uint64_t seed,
                      void* out_host) {
    if (!m || !aux_host || !out_host) return fail("swn_generate_host: null argument");
    if (!m->finalized) return fail("swn_generate_host: call swn_finalize first");
    CUDA_OK(cudaSetDevice(m->device));
    const size_t n_aux = (size_t)B * m->cfg.n_aux * T_f;
    const int64_t n_gen = (n_samples / m->seg) * m->seg;
    const size_t esz = m->cfg.kind == SWN_KIND_CSWNV ? sizeof(float) : sizeof(long long);
    if (m->e2e_aux.ensure(n_aux) || m->e2e_out.ensure((size_t)B * n_gen * esz)) return 1;
    CUDA_OK(cudaMemcpyAsync(m->e2e_aux.p, aux_host, n_aux * sizeof(float), cudaMemcpyHostToDevice, 0));
    if (swn_condition(m, m->e2e_aux.p, B, T_f, nullptr)) return 1;
    swn_gen_args a;
    memset(&a, 0, sizeof(a));
    a.n_samples = n_samples; a.seed = seed; a.out_dev = m->e2e_out.p;
    if (swn_generate(m, &a, nullptr)) return 1;
    CUDA_OK(cudaMemcpyAsync(out_host, m->e2e_out.p, (size_t)B * n_gen * esz, cudaMemcpyDeviceToHost, 0));
    return swn_sync(m, nullptr);
}

int swn_get_info(swn_handle m, swn_info* info) {
    if (!m || !info) return fail("swn_get_info: null argument");
    *info = m->info;
    return 0;
}

int swn_philox_uniform(uint64_t seed, int32_t kind, int32_t n_quantize, int64_t b, int64_t n0, int64_t n, float* out_host) {
    if (!out_host) return fail("swn_philox_uniform: null argument");
    if (kind == SWN_KIND_CSWNV) {
        for (int64_t i = 0; i < n; ++i) out_host[i] = swn::philox_uniform(seed, b, n0 + i, 0);
    } else {
        for (int64_t i = 0; i < n; ++i)
            for (int r = 0; r < n_quantize; ++r) out_host[i * n_quantize + r] = swn::philox_uniform(seed, b, n0 + i, r);
    }
    return 0;
}

int swn_measure_smem_bandwidth(int device, double* gb_per_s) {
    if (!gb_per_s) return fail("swn_measure_smem_bandwidth: null argument");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("swn_measure_smem_bandwidth: device is not sm_100");
    CUDA_OK(swn::measure_smem_bandwidth(prop.multiProcessorCount, gb_per_s, nullptr));
    return 0;
}

int swn_noise_shaping(int device, const swn_ns_args* a, void* stream) {
    if (!a || !a->x_dev || !a->y_dev || !a->n_samples) return fail("swn_noise_shaping: null argument");
    const bool fir_only = a->mcep_host == nullptr;               // low_cut_filter alone (feature_extract.py:57-76)
    if (fir_only && a->fir_taps < 1) return fail("swn_noise_shaping: neither a mel-cepstrum nor FIR taps were given");
    if (a->B < 1 || a->stride < 1) return fail("swn_noise_shaping: B=%d stride=%lld", a->B, (long long)a->stride);
    if (!fir_only) {
        if (a->order < 2 || a->order > 62) return fail("swn_noise_shaping: order=%d is outside 2..62", a->order);
        if (a->pade != 4 && a->pade != 5) return fail("swn_noise_shaping: pade=%d (4 or 5)", a->pade);
        if (a->hop < 1 || a->mcep_frames < 1 || (a->mcep_utts != 1 && a->mcep_utts != a->B))
            return fail("swn_noise_shaping: hop=%d mcep_frames=%d mcep_utts=%d (1 or B=%d)", a->hop, a->mcep_frames, a->mcep_utts, a->B);
    }
    if (a->fir_taps > 0 && (!a->fir_host || (!fir_only && !a->tmp_dev)))
        return fail("swn_noise_shaping: fir_taps=%d needs fir_host and tmp_dev", a->fir_taps);
    CUDA_OK(cudaSetDevice(device));
    {
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) return fail("swn_noise_shaping: device %d is sm_%d%d, this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int B = a->B, m = a->order;
    // mc2b on the host (SPTK mc2b.c: b(m) = c(m), b(i) = c(i) - alpha b(i+1)): a few hundred values
    std::vector<double> b(fir_only ? 0 : (size_t)a->mcep_utts * a->mcep_frames * (m + 1));
    for (size_t r = 0; !fir_only && r < (size_t)a->mcep_utts * a->mcep_frames; ++r) {
        const double* c = a->mcep_host + r * (m + 1);
        double* o = b.data() + r * (m + 1);
        o[m] = c[m];
        for (int i = m - 1; i >= 0; --i) o[i] = c[i] - a->alpha * o[i + 1];
        for (int i = 0; i <= m; ++i)
            if (!(o[i] == o[i]) || o[i] > 1e300 || o[i] < -1e300) return fail("swn_noise_shaping: non-finite filter coefficient");
    }
    std::vector<long long> ns(B);
    std::vector<int> nf(B);
    long long n_max = 0;
    for (int u = 0; u < B; ++u) {
        if (a->n_samples[u] < 0 || a->n_samples[u] > a->stride) return fail("swn_noise_shaping: n_samples[%d]=%lld", u, (long long)a->n_samples[u]);
        ns[u] = a->n_samples[u]; n_max = std::max(n_max, ns[u]);
        long long f = a->n_frames ? a->n_frames[u] : a->mcep_frames;
        if (a->mcep_frames > 1 && f > a->mcep_frames) f = a->mcep_frames;
        nf[u] = (int)std::max<long long>(0, std::min<long long>(f, 0x7fffffff));
    }
    DevBuf<double> b_dev, h_dev;
    DevBuf<long long> ns_dev;
    DevBuf<int> nf_dev;
    int rc = 0;
    if (upload(b_dev, b) || upload(ns_dev, ns) || upload(nf_dev, nf)) rc = 1;
    if (!rc && a->fir_taps > 0) {
        std::vector<double> h(a->fir_host, a->fir_host + a->fir_taps);
        if (upload(h_dev, h)) rc = 1;
    }
    cudaError_t e = cudaSuccess;
    if (!rc && fir_only) {
        e = swn::launch_fir(a->x_dev, a->y_dev, a->stride, ns_dev.p, n_max, h_dev.p, a->fir_taps, B, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    } else if (!rc) {
        double* mid = a->fir_taps > 0 ? a->tmp_dev : a->y_dev;
        e = swn::launch_mlsa(a->x_dev, mid, a->stride, ns_dev.p, nf_dev.p, b_dev.p, a->mcep_utts, a->mcep_frames, m, a->pade, a->alpha,
                             a->hop, B, st);
        if (e == cudaSuccess && a->fir_taps > 0)
            e = swn::launch_fir(mid, a->y_dev, a->stride, ns_dev.p, n_max, h_dev.p, a->fir_taps, B, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // the coefficient buffers are released below
    }
    b_dev.release(); h_dev.release(); ns_dev.release(); nf_dev.release();
    if (rc) return 1;
    if (e != cudaSuccess) return fail("swn_noise_shaping: %s", cudaGetErrorString(e));
    return 0;
}

int swn_destroy(swn_handle m) {
    if (!m) return 0;
    cudaSetDevice(m->device);
    m->wbuf.release(); m->pconst.release(); m->gain.release(); m->in_tab.release(); m->in_bias.release();
    m->xa_tab.release(); m->wp.release(); m->act0.release(); m->act1.release(); m->ptab.release();
    m->desc.release(); m->ll.release(); m->ring_scratch.release(); m->status.release();
    m->e2e_aux.release(); m->e2e_out.release(); m->tf_out.release();
    m->wbuf3.release(); m->tiles3.release(); m->skip3.release(); m->ring3.release(); m->hist3.release(); m->slot3.release();
    m->cnt3.release(); m->desc3.release();
    m->wbuf2.release(); m->ring2.release(); m->cdesc.release(); m->hdesc.release(); m->ll2.release();
    for (auto& cw : m->aux_convs) { cw.w_t.release(); cw.b.release(); }
    for (int i = 0; i < 4; ++i) if (m->ev[i]) cudaEventDestroy(m->ev[i]);
    delete m;
    return 0;
}

}  // extern "C"
