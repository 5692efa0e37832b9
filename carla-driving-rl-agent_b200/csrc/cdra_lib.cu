// libcdra: C ABI (include/cdra.h) over the hand-written sm_100a kernels.
#include <cstring>
#include <string>
#include <cmath>
#include "plan.h"
#include "tower_run.cuh"
#include "tower_opt.cuh"
#include "v2_run.cuh"
#include "tail.cuh"
#include "ppo.cuh"
#include "augment.cuh"

using namespace cdra;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#ifdef CDRA_EMU
static int check_launch(const char*) { return CDRA_OK; }
static void zero_async(void* p, size_t n, cudaStream_t) { memset(p, 0, n); }
#else
static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CDRA_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return CDRA_OK;
}
static void zero_async(void* p, size_t n, cudaStream_t s) { cudaMemsetAsync(p, 0, n, s); }
#endif

struct cdra_plan { Plan* p; };

// ----------------------------------------------------------------------------------------------- launch accounting
#include <atomic>
#include <map>
#include <mutex>
namespace {
struct ProfEntry { long long count = 0; double ms = 0.0, bytes = 0.0; };
std::atomic<long long> g_launches{0};
bool g_prof = false;
double g_pending_bytes = 0.0;
std::map<const void*, ProfEntry> g_entries;
#ifndef CDRA_EMU
cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
const void* g_cur = nullptr;
const void* g_cur_dbg = nullptr;
#endif
}
namespace cdra {
void prof_bytes(double b) { g_pending_bytes = b; }
#ifndef CDRA_EMU
void prof_pre(const void* func, cudaStream_t stream) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    g_cur_dbg = func;
    if (!g_prof) return;
    if (!g_ev0) { cudaEventCreate(&g_ev0); cudaEventCreate(&g_ev1); }
    g_cur = func;
    cudaEventRecord(g_ev0, stream);
}
void prof_post(cudaStream_t stream) {
    static const bool dbg = getenv("CDRA_DEBUG_SYNC") != nullptr;      // debugging aid: fail at the offending launch
    if (dbg) {
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            const char* name = nullptr;
            if (g_cur_dbg) cudaFuncGetName(&name, g_cur_dbg);
            fprintf(stderr, "libcdra: launch %lld (%s) failed: %s\n", (long long)g_launches.load(), name ? name : "?", cudaGetErrorString(e));
        }
    }
    if (!g_prof) { g_pending_bytes = 0.0; return; }
    cudaEventRecord(g_ev1, stream);
    cudaEventSynchronize(g_ev1);
    float ms = 0.f; cudaEventElapsedTime(&ms, g_ev0, g_ev1);
    ProfEntry& e = g_entries[g_cur];
    e.count++; e.ms += ms; e.bytes += g_pending_bytes; g_pending_bytes = 0.0;
}
#endif
}

#ifndef CDRA_EMU
bool cdra::pdl_enabled() { static const bool v = getenv("CDRA_NO_PDL") == nullptr; return v; }
#endif

// ----------------------------------------------------------------------------------------------- small launch helpers
static thread_local bool t_gemm_tf32 = false;     // set by every ABI entry: bf16 perf mode runs the dense GEMMs on tensor cores
static void gemm(cudaStream_t st, bool ta, bool tb, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                 const float* bias, int M, int N, int K, bool accumulate) {
    GemmArgs g{A, lda, B, ldb, C, ldc, bias, M, N, K, accumulate ? 1 : 0};
    dim3 grid(cdiv(M, kGT), cdiv(N, kGT));
#ifndef CDRA_EMU
    if (t_gemm_tf32) {
        TGemmArgs t{g, (((uintptr_t)A & 15) == 0 && lda % 4 == 0) ? 1 : 0, (((uintptr_t)B & 15) == 0 && ldb % 4 == 0) ? 1 : 0};
        if (!ta && !tb) { auto k = tgemm_kernel<false, false>; CDRA_LAUNCH(k, grid, dim3(256), 0, st, t); }
        else if (!ta && tb) { auto k = tgemm_kernel<false, true>; CDRA_LAUNCH(k, grid, dim3(256), 0, st, t); }
        else if (ta && !tb) { auto k = tgemm_kernel<true, false>; CDRA_LAUNCH(k, grid, dim3(256), 0, st, t); }
        else { auto k = tgemm_kernel<true, true>; CDRA_LAUNCH(k, grid, dim3(256), 0, st, t); }
        return;
    }
#endif
    if (!ta && !tb) { auto k = sgemm_kernel<false, false>; CDRA_LAUNCH(k, grid, dim3(256), 0, st, g); }
    else if (!ta && tb) { auto k = sgemm_kernel<false, true>; CDRA_LAUNCH(k, grid, dim3(256), 0, st, g); }
    else if (ta && !tb) { auto k = sgemm_kernel<true, false>; CDRA_LAUNCH(k, grid, dim3(256), 0, st, g); }
    else { auto k = sgemm_kernel<true, true>; CDRA_LAUNCH(k, grid, dim3(256), 0, st, g); }
}
static void colsum(cudaStream_t st, const float* X, int ldx, int M, int N, float* out, bool accumulate) {
    ColsumArgs a{X, ldx, M, N, out, accumulate ? 1 : 0};
    CDRA_LAUNCH(colsum_kernel, dim3(cdiv(N, 32)), dim3(256), 0, st, a);
}
static inline float* F(char* ws, size_t off) { return (float*)(ws + off); }
static size_t named_off(const Plan& p, const char* name) { return p.named.at(name).first; }

// ----------------------------------------------------------------------------------------------- dynamics tail (forward)
static void feat_args(const Plan& p, char* ws, const float* params, float* state, float* grads,
                      const float* const x[3], int training, FeatArgs3& aa) {
    aa.B = p.B; aa.training = training;
    for (int m = 0; m < 3; ++m) {
        const FeatSpec& f = p.feats[m];
        FeatArgs& a = aa.m[m];
        a.x = x[m]; a.d = f.d;
        a.w1 = params + f.d1.w; a.b1 = params + f.d1.b; a.g1 = params + f.d1.g; a.be1 = params + f.d1.be;
        a.w2 = params + f.d2.w; a.b2 = params + f.d2.b; a.g2 = params + f.d2.g; a.be2 = params + f.d2.be;
        a.mm1 = state ? state + f.d1.mm : nullptr; a.mv1 = state ? state + f.d1.mv : nullptr;
        a.mm2 = state ? state + f.d2.mm : nullptr; a.mv2 = state ? state + f.d2.mv : nullptr;
        a.h1 = F(ws, f.h1); a.h2 = F(ws, f.h2); a.n1 = F(ws, f.n1); a.out = F(ws, f.out);
        a.st1 = (float2*)(ws + f.st1); a.st2 = (float2*)(ws + f.st2);
        a.dout = F(ws, f.dbuf1); a.dtmp = F(ws, f.dbuf2);
        if (grads) {
            a.dw1 = grads + f.d1.w; a.db1 = grads + f.d1.b; a.dg1 = grads + f.d1.g; a.dbe1 = grads + f.d1.be;
            a.dw2 = grads + f.d2.w; a.db2 = grads + f.d2.b; a.dg2 = grads + f.d2.g; a.dbe2 = grads + f.d2.be;
        } else { a.dw1 = a.db1 = a.dg1 = a.dbe1 = a.dw2 = a.db2 = a.dg2 = a.dbe2 = nullptr; }
    }
}

// ---- side stream: tower-independent work runs next to the image tower; fork = "side may start after everything
// enqueued on the caller's stream so far", join = "the caller's stream continues after the side work"
#ifndef CDRA_EMU
static cudaStream_t side_stream(const RunCtx& c) {
    static const bool off = getenv("CDRA_NO_SIDE") != nullptr;
    if (off || g_prof) return c.stream;
    const Plan& p = *c.p;
    if (!p.side_stream) {
        cudaStream_t s; cudaEvent_t e0, e1, e2;
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&e0, cudaEventDisableTiming); cudaEventCreateWithFlags(&e1, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e2, cudaEventDisableTiming);
        p.side_stream = s; p.ev_fork = e0; p.ev_join = e1; p.ev_prep = e2;
    }
    return (cudaStream_t)p.side_stream;
}
static void side_fork(const RunCtx& c, cudaStream_t sd) {
    if (sd == c.stream) return;
    cudaEventRecord((cudaEvent_t)c.p->ev_fork, c.stream);
    cudaStreamWaitEvent(sd, (cudaEvent_t)c.p->ev_fork, 0);
}
static void side_join(const RunCtx& c, cudaStream_t sd) {
    if (sd == c.stream) return;
    cudaEventRecord((cudaEvent_t)c.p->ev_join, sd);
    cudaStreamWaitEvent(c.stream, (cudaEvent_t)c.p->ev_join, 0);
}
static void side_destroy(Plan* p) {
    if (p->side_stream) {
        cudaStreamDestroy((cudaStream_t)p->side_stream);
        cudaEventDestroy((cudaEvent_t)p->ev_fork); cudaEventDestroy((cudaEvent_t)p->ev_join); cudaEventDestroy((cudaEvent_t)p->ev_prep);
        p->side_stream = p->ev_fork = p->ev_join = p->ev_prep = nullptr;
    }
}
#else
static cudaStream_t side_stream(const RunCtx& c) { return c.stream; }
static void side_fork(const RunCtx&, cudaStream_t) {}
static void side_join(const RunCtx&, cudaStream_t) {}
static void side_destroy(Plan*) {}
#endif

// which = 1: feature MLPs + their three GRUs (independent of the image tower);  2: image GRU + trunk;  3: both
static void tail_forward(const RunCtx& c, const float* road, const float* vehicle, const float* nav, float* out512,
                         cudaStream_t st, int which) {
    const Plan& p = *c.p; const int B = p.B;
    if (which & 1) {   // feature MLPs                                    core/networks.py:41-43
        FeatArgs3 aa; const float* x[3] = {road, vehicle, nav};
        feat_args(p, c.ws, c.params, c.state, nullptr, x, c.training, aa);
        CDRA_LAUNCH(featnet_fwd_kernel, dim3(3), dim3(kFeatThreads), 0, st, aa);
    }
    int col = 0, gi = 0;
    for (const GruSpec& g : p.grus) {                                    // core/networks.py:46-50
        const int u = g.units, u3 = 3 * u;
        if (!(which & (gi++ == 0 ? 2 : 1))) { col += u; continue; }
        const float* K = c.params + g.k; const float* R = c.params + g.r; const float* b = c.params + g.b;
        gemm(st, false, false, F(c.ws, g.x_in), g.din, K, u3, F(c.ws, g.xp), u3, b, 4 * B, u3, g.din, false);
        for (int t = 0; t < kT; ++t) {
            GruGateArgs a; memset(&a, 0, sizeof a);
            a.B = B; a.u = u;
            a.xp = F(c.ws, g.xp) + (size_t)t * B * u3; a.hp = F(c.ws, g.hp) + (size_t)t * B * u3; a.b1 = b + u3;
            if (t > 0) {
                a.hprev = F(c.ws, g.hs) + (size_t)(t - 1) * B * u; a.ldhp = u;
                gemm(st, false, false, a.hprev, u, R, u3, a.hp, u3, b + u3, B, u3, u, false);
            }
            if (t == kT - 1) { a.hout = F(c.ws, p.dyn_in) + col; a.ldho = 352; }
            else { a.hout = F(c.ws, g.hs) + (size_t)t * B * u; a.ldho = u; }
            CDRA_LAUNCH(gru_gate_fwd_kernel, dim3(cdiv((long long)B * u, 256)), dim3(256), 0, st, a);
        }
        col += u;
    }
    if (which & 2) {   // linear_combination: BN(352) -> Dense 512       core/networks.py:24-30,53-55
        Bn1dArgs a; memset(&a, 0, sizeof a);
        a.x = F(c.ws, p.dyn_in); a.ldx = 352; a.y = F(c.ws, p.trunk_n); a.ldy = 352;
        a.stat = (float2*)(c.ws + p.trunk_stat); a.gamma = c.params + p.trunk_g; a.beta = c.params + p.trunk_be;
        a.mov_mean = c.state ? c.state + p.trunk_mm : nullptr; a.mov_var = c.state ? c.state + p.trunk_mv : nullptr;
        a.B = B; a.C = 352; a.training = c.training;
        CDRA_LAUNCH(bn1d_fwd_kernel, dim3(cdiv(352, 32)), dim3(256), 0, st, a);
        gemm(st, false, false, F(c.ws, p.trunk_n), 352, c.params + p.trunk_w, 512, out512, 512, c.params + p.trunk_b, B, 512, 352, false);
    }
}

// ----------------------------------------------------------------------------------------------- inference-mode tables
static void eval_affine(const RunCtx& c) {
    const Plan& p = *c.p;
    EvalAffArgs a; a.n = 0;
    auto flush = [&]() {
        if (a.n) { CDRA_LAUNCH(eval_affine_kernel, dim3(a.n), dim3(256), 0, c.stream, a); a.n = 0; }
    };
    auto add = [&](const BnConv& l, const WsTensor& dst, ColMap cm) {
        EvalAffLayer& L = a.l[a.n++];
        L.gamma = c.params + l.g; L.beta = c.params + l.be; L.mm = c.state + l.mm; L.mv = c.state + l.mv;
        L.aff = (float2*)(c.ws + dst.aff); L.bnp = (float2*)(c.ws + dst.bnp); L.ld = dst.C; L.cm = cm;
        if (a.n == kEvalMax) flush();
    };
    add(p.stem, p.tensors[p.t_stem], ColMap{kStemC, 0, 0, 0});
    for (const Unit& u : p.units) {
        const int sc = u.stride == 2 ? u.cin : u.cin / 2;
        add(u.pw1, p.tensors[u.t_r1], ColMap{u.half, 0, 0, 0});
        add(u.dw, p.tensors[u.t_r2], ColMap{u.half, 0, 0, 0});
        add(u.pw2, p.tensors[u.t_out], ColMap{u.c - sc, 1, sc / 2, u.half});
        if (u.stride == 2) {
            add(u.scdw, p.tensors[u.t_rs], ColMap{sc, 0, 0, 0});
            add(u.scpw, p.tensors[u.t_out], ColMap{sc, 1, 0, u.half});
        }
    }
    add(p.head, p.tensors[p.t_head], ColMap{p.head.N, 0, 0, 0});
    flush();
}

// ----------------------------------------------------------------------------------------------- tower backward
template <typename T>
static void launch_bstat(const RunCtx& c, const WsTensor& t, int coff, int C, bool clamp) {
    BstatArgs<T> a;
    a.dA = (const T*)(c.ws + t.grad); a.R = (const T*)(c.ws + t.data); a.ld = t.C; a.coff = coff; a.C = C; a.Rt = t.Rt;
    a.clamp = clamp ? 1 : 0; a.tb = tables_of(c, t); a.counter = nullptr;    // sums go straight to replica 0 (no fold, no ticket)
    int rows = 16384 / C; if (rows < 64) rows = 64;
    a.rows_per_block = rows;
    prof_bytes(4.0 * t.Rt * C * 2 * sizeof(T));                // read dA and R once
    auto k = bstat_kernel<T>;
    CDRA_LAUNCH(k, dim3(cdiv(t.Rt, rows), kT), dim3(256), 0, c.stream, a);
    // gradient wrt the activated value -> gradient wrt the raw conv output, in place (BN + ReLU6 backward);
    // every conv-backward kernel downstream consumes dR as a plain matrix
    prof_bytes(4.0 * t.Rt * C * 3 * sizeof(T));
    auto k2 = dr_kernel<T>;
    CDRA_LAUNCH(k2, dim3(cdiv(t.Rt, rows), kT), dim3(256), 0, c.stream, a);
}

template <typename T>
static void launch_pw_bwd(const RunCtx& c, const BnConv& l, const WsTensor& out, const ColMap& cm,
                          const ActView& in, int Rt, T* dx, int ldx, int coffx, bool accumulate, bool need_dx) {
    PwBwdArgs<T> a;
    a.out = (const T*)(c.ws + out.data); a.dout = (const T*)(c.ws + out.grad); a.ldo = out.C; a.cm = cm;
    a.tb = tables_of(c, out); a.clamp = 1;
    a.in = in; a.K = l.K; a.Rt = Rt; a.w = c.params + l.w;
    a.dx = dx; a.ldx = ldx; a.coffx = coffx; a.accumulate = accumulate ? 1 : 0;
    a.dw = c.grads + l.w; a.db = c.grads + l.b; a.dgamma = c.grads + l.g; a.dbeta = c.grads + l.be;
    int splits = (int)cdiv(Rt, 2048); if (splits < 1) splits = 1; if (splits > 64) splits = 64;
    a.row_splits = splits; a.partials = nullptr;
#ifndef CDRA_EMU
    if constexpr (std::is_same<T, bf16>::value) if (use_mma()) {               // tensor-core path (pw_mma.cuh)
        if (need_dx) {
            PwMmaBwdArgs pa; pa.a = a; pa.wn = (const bf16*)(c.ws + l.wn); pa.Np = l.Np;
            prof_bytes(4.0 * Rt * (2.0 * cm.n + l.K) * sizeof(T));
            auto k64 = pw_dgrad_mma_kernel<64>;
            static bool carve = (cudaFuncSetAttribute(k64, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared),
                                 cudaFuncSetAttribute(pw_wgrad_mma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true);
            (void)carve;
            CDRA_LAUNCH(k64, dim3(cdiv(Rt, kMmTM), kT, cdiv(l.K, 64)), dim3(256), 0, c.stream, pa);
        }
        const int kt = (int)cdiv(l.K + 1, kWgKT), nt = (int)cdiv(cm.n, kWgNT);
        int sp = 888 / (kt * nt * kT); if (sp < 1) sp = 1;             // ~6 CTAs per SM in total
        const int max_sp = (int)cdiv(Rt, 64); if (sp > max_sp) sp = max_sp;
        a.row_splits = sp;
        a.partials = F(c.ws, named_off(*c.p, "wgrad.partials"));      // kt*nt*4*sp <= 1024 tiles
        prof_bytes(4.0 * Rt * (2.0 * cm.n + l.K) * sizeof(T));
        CDRA_LAUNCH(pw_wgrad_mma_kernel, dim3(kt, nt, kT * sp), dim3(256), 0, c.stream, a);
        PwWgReduceArgs ra{a, kt, nt, kT * sp};
        CDRA_LAUNCH(pw_wgrad_reduce_kernel, dim3(kt, nt, kWgKT * kWgNT / 32), dim3(256), 0, c.stream, ra);
        return;
    }
#endif
    if (need_dx) {
        prof_bytes(4.0 * Rt * (2.0 * cm.n + l.K) * sizeof(T));  // read dA, R; write dX
        auto k = pw_dgrad_kernel<T>;
        CDRA_LAUNCH(k, dim3(cdiv(Rt, kPwTM), kT, cdiv(l.K, kPwTN)), dim3(256), 0, c.stream, a);
    }
    prof_bytes(4.0 * Rt * (2.0 * cm.n + l.K) * sizeof(T));      // read dA, R, X
    auto k2 = pw_wgrad_kernel<T>;
    CDRA_LAUNCH(k2, dim3(cdiv(l.K + 1, kPwTM), cdiv(cm.n, kPwTN), kT * splits), dim3(256), 0, c.stream, a);
}

template <typename T>
static void launch_dw_bwd(const RunCtx& c, const BnConv& l, const WsTensor& out, const ActView& in, const Unit& u, int C,
                          T* dx, int ldx, int coffx, bool accumulate) {
    DwBwdArgs<T> a;
    a.out = (const T*)(c.ws + out.data); a.dout = (const T*)(c.ws + out.grad); a.tb = tables_of(c, out);
    a.in = in; a.B = c.p->B; a.Hi = u.Hi; a.Wi = u.Wi; a.Ho = u.Ho; a.Wo = u.Wo; a.C = C; a.stride = u.stride;
    a.pad_t = u.pad_t; a.pad_l = u.pad_l; a.w = c.params + l.w;
    a.dx = dx; a.ldx = ldx; a.coffx = coffx; a.accumulate = accumulate ? 1 : 0;
    a.dw = c.grads + l.w; a.db = c.grads + l.b; a.dgamma = c.grads + l.g; a.dbeta = c.grads + l.be;
    int lanes_c = ((C / 2) + 31) & ~31; if (lanes_c > 256) lanes_c = 256;
    a.ppb = 16 * (256 / lanes_c); a.ppb_w = 64 * (256 / lanes_c);
    prof_bytes(4.0 * a.B * (2.0 * u.Ho * u.Wo + (double)u.Hi * u.Wi) * C * sizeof(T));
    const int lanes_r = 256 / lanes_c;
    if (u.stride == 1) {
        auto k1 = dw_dgrad_row_kernel<T>;
        CDRA_LAUNCH(k1, dim3(cdiv((long long)a.B * u.Hi, lanes_r), kT), dim3(256), 0, c.stream, a);
    } else {
        auto k1 = dw_dgrad_kernel<T>;
        CDRA_LAUNCH(k1, dim3(cdiv((long long)a.B * u.Hi * u.Wi, a.ppb), kT), dim3(256), 0, c.stream, a);
    }
    prof_bytes(4.0 * a.B * (2.0 * u.Ho * u.Wo + (double)u.Hi * u.Wi) * C * sizeof(T));
    unsigned wg_blocks = cdiv((long long)a.B * u.Ho, lanes_r); if (wg_blocks > 148) wg_blocks = 148;   // persistent: 4 CTAs / SM over the 4 slices
    if (u.stride == 1) { auto k2 = dw_wgrad_row_kernel<T, 1>; CDRA_LAUNCH(k2, dim3(wg_blocks, kT), dim3(256), 0, c.stream, a); }
    else { auto k2 = dw_wgrad_row_kernel<T, 2>; CDRA_LAUNCH(k2, dim3(wg_blocks, kT), dim3(256), 0, c.stream, a); }
}

template <typename T, typename TIn>
static void tower_backward(const RunCtx& c, const TIn* image, bool stem_only = false) {
    const Plan& p = *c.p; const int B = p.B;
    const WsTensor& th = p.tensors[p.t_head];
    if (!stem_only) {   // GAP backward, head conv backward
        GapBwdArgs<T> a{F(c.ws, p.dgap), (T*)(c.ws + th.grad), B, th.H * th.W, th.C};
        auto k = gap_bwd_kernel<T>;
        CDRA_LAUNCH(k, dim3(cdiv((long long)B * th.H * th.W * th.C, 256), kT), dim3(256), 0, c.stream, a);
        launch_bstat<T>(c, th, 0, th.C, true);
        const WsTensor& tin = p.tensors[p.units.back().t_out];
        launch_pw_bwd<T>(c, p.head, th, ColMap{p.head.N, 0, 0, 0}, view_of(c, tin, 0, true), tin.Rt,
                         (T*)(c.ws + tin.grad), tin.C, 0, false, true);
    }
    for (int ui = (int)p.units.size() - 1; ui >= 0 && !stem_only; --ui) {
        const Unit& u = p.units[ui];
        const WsTensor& tin = p.tensors[u.t_in];
        const WsTensor& r1 = p.tensors[u.t_r1];
        const WsTensor& r2 = p.tensors[u.t_r2];
        const WsTensor& out = p.tensors[u.t_out];
        const bool in_clamp = tin.tables;
        const int sc = u.stride == 2 ? u.cin : u.cin / 2;
        T* dxin = (T*)(c.ws + tin.grad);
        if (u.stride == 1) {
            // the shortcut half first: the in-place dA -> dR pass below also rewrites the pass-through channels of
            // the unit-output gradient (harmless once they have been copied out)
            PassBwdArgs<T> a{(const T*)(c.ws + out.grad), out.C, dxin, tin.C, u.half, out.Rt};
            auto k = pass_bwd_kernel<T>;
            CDRA_LAUNCH(k, dim3(cdiv((long long)out.Rt * (u.half / 2), 256), kT), dim3(256), 0, c.stream, a);
        }
        launch_bstat<T>(c, out, 0, out.C, true);
        // branch
        launch_pw_bwd<T>(c, u.pw2, out, ColMap{u.c - sc, 1, sc / 2, u.half}, view_of(c, r2, 0, false), r2.Rt,
                         (T*)(c.ws + r2.grad), r2.C, 0, false, true);
        launch_bstat<T>(c, r2, 0, r2.C, false);
        launch_dw_bwd<T>(c, u.dw, r2, view_of(c, r1, 0, true), u, u.half, (T*)(c.ws + r1.grad), r1.C, 0, false);
        launch_bstat<T>(c, r1, 0, r1.C, true);
        const int xoff = u.stride == 2 ? 0 : u.cin / 2;
        launch_pw_bwd<T>(c, u.pw1, r1, ColMap{u.half, 0, 0, 0}, view_of(c, tin, xoff, in_clamp), tin.Rt,
                         dxin, tin.C, xoff, false, true);
        if (u.stride == 2) {
            const WsTensor& rs = p.tensors[u.t_rs];
            launch_pw_bwd<T>(c, u.scpw, out, ColMap{sc, 1, 0, u.half}, view_of(c, rs, 0, false), rs.Rt,
                             (T*)(c.ws + rs.grad), rs.C, 0, false, true);
            launch_bstat<T>(c, rs, 0, rs.C, false);
            launch_dw_bwd<T>(c, u.scdw, rs, view_of(c, tin, 0, in_clamp), u, sc, dxin, tin.C, 0, true);
        }
    }
    {   // maxpool + stem
        const WsTensor& ts = p.tensors[p.t_stem];
        const WsTensor& tp = p.tensors[p.t_pool];
        PoolBwd2Args<T> a;
        a.in = view_of(c, ts, 0, true); a.pool = (const T*)(c.ws + tp.data); a.dpool = (const T*)(c.ws + tp.grad);
        a.dstem = (T*)(c.ws + ts.grad);
        a.B = B; a.Hi = p.Hs; a.Wi = p.Ws; a.Ho = p.Hp; a.Wo = p.Wp; a.C = kStemC; a.pad_t = p.pool_pad_t; a.pad_l = p.pool_pad_l;
        prof_bytes(4.0 * B * ((double)p.Hs * p.Ws * 2 + (double)p.Hp * p.Wp * 2) * kStemC * sizeof(T));
        auto k = pool_bwd2_kernel<T>;
        CDRA_LAUNCH(k, dim3(p.Hs, B, kT), dim3(256), 0, c.stream, a);
        launch_bstat<T>(c, ts, 0, kStemC, true);
        StemBwdArgs<T, TIn> s;
        s.img = image; s.B = B; s.H = p.H; s.W = p.W; s.Ho = p.Hs; s.Wo = p.Ws;
        s.out = (const T*)(c.ws + ts.data); s.dout = (const T*)(c.ws + ts.grad); s.tb = tables_of(c, ts);
        s.dw = c.grads + p.stem.w; s.db = c.grads + p.stem.b; s.dgamma = c.grads + p.stem.g; s.dbeta = c.grads + p.stem.be;
        int ppb = 1024; s.pix_per_block = ppb;
        auto k2 = stem_wgrad_kernel<T, TIn>;
        CDRA_LAUNCH(k2, dim3(cdiv(ts.Rt, ppb), kT), dim3(kStemWgThreads), 0, c.stream, s);
    }
}

// ----------------------------------------------------------------------------------------------- dynamics tail (backward)
static void tail_backward(const RunCtx& c, const float* road, const float* vehicle, const float* nav, const float* d_out512) {
    const Plan& p = *c.p; const int B = p.B; cudaStream_t st = c.stream;
    float* dn = F(c.ws, named_off(p, "d.trunk.n"));
    // Dense 512: weight / bias gradients are leaves (side stream), the input gradient continues the chain
    cudaStream_t sd = side_stream(c);
    side_fork(c, sd);
    gemm(sd, true, false, F(c.ws, p.trunk_n), 352, d_out512, 512, c.grads + p.trunk_w, 512, nullptr, 352, 512, B, false);
    colsum(sd, d_out512, 512, B, 512, c.grads + p.trunk_b, false);
    gemm(st, false, true, d_out512, 512, c.params + p.trunk_w, 512, dn, 352, nullptr, B, 352, 512, false);
    {
        Bn1dArgs a; memset(&a, 0, sizeof a);
        a.x = F(c.ws, p.dyn_in); a.ldx = 352; a.y = F(c.ws, p.ddyn_in); a.ldy = 352; a.dy = dn; a.lddy = 352;
        a.stat = (float2*)(c.ws + p.trunk_stat); a.gamma = c.params + p.trunk_g; a.beta = c.params + p.trunk_be;
        a.dgamma = c.grads + p.trunk_g; a.dbeta = c.grads + p.trunk_be; a.B = B; a.C = 352;
        CDRA_LAUNCH(bn1d_bwd_kernel, dim3(cdiv(352, 32)), dim3(256), 0, st, a);
    }
    // Side stream: everything the tower's backward does not wait for -- the three feature GRUs' whole backward chains (their input
    // gradients only feed the feature-MLP backward, a leaf), every leaf parameter gradient and the feature-MLP backward.
    // Only the image GRU's chain (d x_in -> global-average-pool backward -> tower) stays on the caller's stream.
    side_fork(c, sd);                    // d(dyn_in) is final
    int col = 0, gi = 0;
    for (const GruSpec& g : p.grus) {
        const bool image = gi++ == 0;
        cudaStream_t cs = image ? st : sd;
        const int u = g.units, u3 = 3 * u;
        const float* K = c.params + g.k; const float* R = c.params + g.r;
        float* dhbuf = F(c.ws, g.dh);
        for (int t = kT - 1; t >= 0; --t) {
            GruGateArgs a; memset(&a, 0, sizeof a);
            a.B = B; a.u = u;
            a.xp = F(c.ws, g.xp) + (size_t)t * B * u3; a.hp = F(c.ws, g.hp) + (size_t)t * B * u3;
            if (t > 0) { a.hprev = F(c.ws, g.hs) + (size_t)(t - 1) * B * u; a.ldhp = u; }
            if (t == kT - 1) { a.dh = F(c.ws, p.ddyn_in) + col; a.lddh = 352; }
            else { a.dh = dhbuf + (size_t)((t + 1) & 1) * B * u; a.lddh = u; }
            a.dxp = F(c.ws, g.dxp) + (size_t)t * B * u3; a.dhp = F(c.ws, g.dhp) + (size_t)t * B * u3;
            if (t > 0) { a.dhprev = dhbuf + (size_t)(t & 1) * B * u; a.lddhp = u; }
            CDRA_LAUNCH(gru_gate_bwd_kernel, dim3(cdiv((long long)B * u, 256)), dim3(256), 0, cs, a);
            if (t > 0)      // dh_{t-1} += dhp_t R^T
                gemm(cs, false, true, a.dhp, u3, R, u3, a.dhprev, u, nullptr, B, u, u3, true);
        }
        gemm(cs, false, true, F(c.ws, g.dxp), u3, K, u3, F(c.ws, g.dx_in), g.din, nullptr, 4 * B, g.din, u3, false);
        // parameter gradients (side stream; the image GRU's wait for its chain on the caller's stream)
        if (image) side_fork(c, sd);
        gemm(sd, true, false, F(c.ws, g.hs), u, F(c.ws, g.dhp) + (size_t)B * u3, u3, c.grads + g.r, u3, nullptr, u, u3, 3 * B, false);
        colsum(sd, F(c.ws, g.dxp), u3, 4 * B, u3, c.grads + g.b, false);
        colsum(sd, F(c.ws, g.dhp), u3, 4 * B, u3, c.grads + g.b + u3, false);
        gemm(sd, true, false, F(c.ws, g.x_in), g.din, F(c.ws, g.dxp), u3, c.grads + g.k, u3, nullptr, g.din, u3, 4 * B, false);
        col += u;
    }
    {   // feature-MLP backward: after the feature GRUs' d x_in (same stream)
        FeatArgs3 aa; const float* x[3] = {road, vehicle, nav};
        feat_args(p, c.ws, c.params, nullptr, c.grads, x, 1, aa);
        CDRA_LAUNCH(featnet_bwd_kernel, dim3(3), dim3(kFeatThreads), 0, sd, aa);
    }
}

// ----------------------------------------------------------------------------------------------- heads
static int run_head(bool policy, cdra_plan_t* plan, const float* params, float* state, const float* x512,
                    const float* actions, const float* actions_jac, const float* logp_old, const float* adv, const float* returns_be,
                    const float* true_speed, const float* true_sim, float clip, float ent_coef, int training,
                    float grad_scale, float* scalars, float* head_out, float* d_x512, float* grads, char* ws, cudaStream_t st) {
    const Plan& p = *plan->p; const HeadSpec& h = policy ? p.policy : p.value; const int B = p.B;
    t_gemm_tf32 = false;      // the heads are tiny and feed the loss directly: fp32 GEMMs in both modes (PPO loss rtol 1e-4 given the trunk output)
    float* n1 = F(ws, named_off(p, "head.n1")); float* pre1 = F(ws, named_off(p, "head.pre1")); float* a1 = F(ws, named_off(p, "head.a1"));
    float* n2 = F(ws, named_off(p, "head.n2")); float* pre2 = F(ws, named_off(p, "head.pre2")); float* a2 = F(ws, named_off(p, "head.a2"));
    float2* st1 = (float2*)(ws + named_off(p, "head.st1")); float2* st2 = (float2*)(ws + named_off(p, "head.st2"));
    float* da2 = F(ws, named_off(p, "d.head.a2")); float* dpre2 = F(ws, named_off(p, "d.head.pre2")); float* dn2 = F(ws, named_off(p, "d.head.n2"));
    float* da1 = F(ws, named_off(p, "d.head.a1")); float* dpre1 = F(ws, named_off(p, "d.head.pre1")); float* dn1 = F(ws, named_off(p, "d.head.n1"));
    double* acc = (double*)(ws + named_off(p, "head.acc"));
    // control_branch forward                                            core/networks.py:59-66
    auto bn_fwd = [&](const float* x, float* y, float2* stt, int C, int64_t g, int64_t be, int64_t mm, int64_t mv) {
        Bn1dArgs a; memset(&a, 0, sizeof a);
        a.x = x; a.ldx = C; a.y = y; a.ldy = C; a.stat = stt; a.gamma = params + g; a.beta = params + be;
        a.mov_mean = state ? state + mm : nullptr; a.mov_var = state ? state + mv : nullptr; a.B = B; a.C = C; a.training = training;
        CDRA_LAUNCH(bn1d_fwd_kernel, dim3(cdiv(C, 32)), dim3(256), 0, st, a);
    };
    auto act = [&](bool fwd, const float* x, const float* dy, float* y, long long n) {
        ActArgs a{x, dy, y, n};
        if (fwd) { CDRA_LAUNCH(swish6_fwd_kernel, dim3(cdiv(n, 256)), dim3(256), 0, st, a); }
        else { CDRA_LAUNCH(swish6_bwd_kernel, dim3(cdiv(n, 256)), dim3(256), 0, st, a); }
    };
    bn_fwd(x512, n1, st1, 512, h.bn1_g, h.bn1_be, h.bn1_mm, h.bn1_mv);
    gemm(st, false, false, n1, 512, params + h.d1_w, kHU, pre1, kHU, params + h.d1_b, B, kHU, 512, false);
    act(true, pre1, nullptr, a1, (long long)B * kHU);
    bn_fwd(a1, n2, st2, kHU, h.bn2_g, h.bn2_be, h.bn2_mm, h.bn2_mv);
    gemm(st, false, false, n2, kHU, params + h.d2_w, kHU, pre2, kHU, params + h.d2_b, B, kHU, kHU, false);
    act(true, pre2, nullptr, a2, (long long)B * kHU);
    // heads + loss (+ gradients wrt logits, head weights, a2)
    zero_async(acc, 32 * sizeof(double), st);
    if (grads) zero_async(grads, (size_t)h.params.size * 4, st);
    HeadLossArgs a; memset(&a, 0, sizeof a);
    a.a2 = a2; a.B = B;
    float* gsink = grads ? grads : F(ws, p.scratch);      // forward-only call: gradients land in scratch
    for (int i = 0; i < 4; ++i) {
        a.w[i] = params + h.out_w[i]; a.b[i] = params + h.out_b[i]; a.n[i] = h.out_n[i];
        a.dw[i] = gsink + (grads ? h.out_w[i] : (int64_t)i * 1024); a.db[i] = gsink + (grads ? h.out_b[i] : (int64_t)4096 + i * 8);
    }
    a.actions = actions; a.actions_jac = actions_jac; a.logp_old = logp_old; a.adv = adv; a.returns_be = returns_be;
    a.true_speed = true_speed; a.true_sim = true_sim; a.clip = clip; a.ent_coef = ent_coef; a.grad_scale = grad_scale;
    a.exp_scale = 6.0f; a.acc = acc; a.scalars = scalars; a.head_out = head_out; a.da2 = da2;
    if (policy) { auto k = head_loss_kernel<true>; CDRA_LAUNCH(k, dim3(cdiv(B, kHeadRows)), dim3(256), 0, st, a); }
    else { auto k = head_loss_kernel<false>; CDRA_LAUNCH(k, dim3(cdiv(B, kHeadRows)), dim3(256), 0, st, a); }
    if (!grads) return check_launch("head forward");
    // control_branch backward
    auto bn_bwd = [&](const float* x, const float* dy, float* dx, float2* stt, int C, int64_t g, int64_t be) {
        Bn1dArgs b; memset(&b, 0, sizeof b);
        b.x = x; b.ldx = C; b.y = dx; b.ldy = C; b.dy = dy; b.lddy = C; b.stat = stt; b.gamma = params + g; b.beta = params + be;
        b.dgamma = grads + g; b.dbeta = grads + be; b.B = B; b.C = C;
        CDRA_LAUNCH(bn1d_bwd_kernel, dim3(cdiv(C, 32)), dim3(256), 0, st, b);
    };
    // the Dense weight / bias gradients are leaves: they run on the plan's side stream next to the chain that produces d x512
    RunCtx c{&p, ws, params, state, grads, st, training};
    cudaStream_t sd = side_stream(c);
    act(false, pre2, da2, dpre2, (long long)B * kHU);
    side_fork(c, sd);
    gemm(sd, true, false, n2, kHU, dpre2, kHU, grads + h.d2_w, kHU, nullptr, kHU, kHU, B, false);
    colsum(sd, dpre2, kHU, B, kHU, grads + h.d2_b, false);
    gemm(st, false, true, dpre2, kHU, params + h.d2_w, kHU, dn2, kHU, nullptr, B, kHU, kHU, false);
    bn_bwd(a1, dn2, da1, st2, kHU, h.bn2_g, h.bn2_be);
    act(false, pre1, da1, dpre1, (long long)B * kHU);
    side_fork(c, sd);
    gemm(sd, true, false, n1, 512, dpre1, kHU, grads + h.d1_w, kHU, nullptr, 512, kHU, B, false);
    colsum(sd, dpre1, kHU, B, kHU, grads + h.d1_b, false);
    gemm(st, false, true, dpre1, kHU, params + h.d1_w, kHU, dn1, 512, nullptr, B, 512, kHU, false);
    bn_bwd(x512, dn1, d_x512, st1, 512, h.bn1_g, h.bn1_be);
    side_join(c, sd);
    return check_launch("head fwd/bwd");
}

// ----------------------------------------------------------------------------------------------- C ABI
extern "C" {

const char* cdra_last_error(void) { return g_err.c_str(); }
int cdra_version(void) { return 100; }

int cdra_plan_create(const cdra_config* cfg, cdra_plan_t** out) {
    if (!cfg || !out) return fail(CDRA_ERR_BADARG, "null argument");
    std::string err;
    Plan* p = build_plan(*cfg, err);
    if (!p) return fail(CDRA_ERR_SHAPE, err);
    *out = new cdra_plan{p};
    return CDRA_OK;
}
void cdra_plan_destroy(cdra_plan_t* plan) { if (plan) { side_destroy(plan->p); delete plan->p; delete plan; } }
size_t cdra_plan_workspace_bytes(const cdra_plan_t* plan) { return plan ? plan->p->ws_bytes : 0; }

static const Arena* arena_of(const cdra_plan_t* plan, int which) {
    if (!plan) return nullptr;
    const Plan& p = *plan->p;
    switch (which) {
        case CDRA_ARENA_DYN_PARAMS: return &p.dyn_params;
        case CDRA_ARENA_DYN_STATE: return &p.dyn_state;
        case CDRA_ARENA_POL_PARAMS: return &p.policy.params;
        case CDRA_ARENA_POL_STATE: return &p.policy.state;
        case CDRA_ARENA_VAL_PARAMS: return &p.value.params;
        case CDRA_ARENA_VAL_STATE: return &p.value.state;
    }
    return nullptr;
}
int64_t cdra_arena_size(const cdra_plan_t* plan, int arena) { const Arena* a = arena_of(plan, arena); return a ? a->size : -1; }
int cdra_arena_num_tensors(const cdra_plan_t* plan, int arena) { const Arena* a = arena_of(plan, arena); return a ? (int)a->tensors.size() : -1; }
int cdra_arena_tensor(const cdra_plan_t* plan, int arena, int index, char* name, int name_cap, int64_t* offset,
                      int32_t* ndim, int32_t dims[4]) {
    const Arena* a = arena_of(plan, arena);
    if (!a || index < 0 || index >= (int)a->tensors.size()) return fail(CDRA_ERR_BADARG, "bad arena / index");
    const ArenaTensor& t = a->tensors[index];
    if (name && name_cap > 0) { strncpy(name, t.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (offset) *offset = t.offset;
    if (ndim) *ndim = t.ndim;
    if (dims) for (int i = 0; i < 4; ++i) dims[i] = t.dims[i];
    return CDRA_OK;
}
int cdra_plan_tensor(const cdra_plan_t* plan, const char* name, int64_t* byte_offset, int32_t dims[4], int32_t* elem_size) {
    if (!plan || !name) return fail(CDRA_ERR_BADARG, "null argument");
    const Plan& p = *plan->p;
    std::string n(name);
    bool grad = false;
    if (n.rfind("grad:", 0) == 0) { grad = true; n = n.substr(5); }
    auto it = p.tensor_index.find(n);
    if (it != p.tensor_index.end()) {
        const WsTensor& t = p.tensors[it->second];
        if (byte_offset) *byte_offset = (int64_t)(grad ? t.grad : t.data);
        if (dims) { dims[0] = 4 * p.B; dims[1] = t.H; dims[2] = t.W; dims[3] = t.C; }
        if (elem_size) *elem_size = t.elem;
        return CDRA_OK;
    }
    auto it2 = p.named.find(n);
    if (it2 == p.named.end()) return fail(CDRA_ERR_BADARG, "unknown tensor " + n);
    if (byte_offset) *byte_offset = (int64_t)it2->second.first;
    if (dims) { for (int i = 0; i < 4; ++i) dims[i] = i < (int)it2->second.second.size() ? it2->second.second[i] : 1; }
    if (elem_size) *elem_size = 4;
    return CDRA_OK;
}

int cdra_debug_export(cdra_plan_t* plan, const char* name, void* workspace, float* out, int32_t dims[4], void* stream) {
    if (!plan || !name) return fail(CDRA_ERR_BADARG, "null argument");
#ifndef CDRA_EMU
    if (plan->p->v2.on && v2::export_tensor(*plan->p, name, (char*)workspace, out, dims, (cudaStream_t)stream))
        return check_launch("debug_export");
#endif
    return fail(CDRA_ERR_BADARG, std::string("no exportable tensor ") + name);
}

int cdra_dynamics_forward(cdra_plan_t* plan, const float* params, float* state, const void* image, const float* road,
                          const float* vehicle, const float* navigation, int training, float* out512, void* workspace,
                          void* stream) {
    if (!plan || !params || !image || !road || !vehicle || !navigation || !out512 || !workspace)
        return fail(CDRA_ERR_BADARG, "null argument");
    if (!training && !state) return fail(CDRA_ERR_BADARG, "inference needs the moving statistics");
    const Plan& p = *plan->p;
    RunCtx c{&p, (char*)workspace, params, state, nullptr, (cudaStream_t)stream, training ? 1 : 0};
    t_gemm_tf32 = p.cfg.dtype == CDRA_DTYPE_BF16 && getenv("CDRA_NO_TF32") == nullptr;
    if (training) zero_async(c.ws, p.zero_bytes, c.stream);
    else eval_affine(c);                 // BN affine from the moving statistics (CARLANetwork.dynamics_predict)
    const bool bf = p.cfg.dtype == CDRA_DTYPE_BF16, u8 = p.cfg.image_u8 != 0;
    cudaStream_t sd = side_stream(c);
    side_fork(c, sd);
#ifndef CDRA_EMU
    if (p.v2.on) {               // GEMM operand preparation runs next to the stem (nothing before the first pointwise layer needs it)
        RunCtx cs = c; cs.stream = sd;
        v2::tower_prepare(cs);
        if (sd != c.stream) cudaEventRecord((cudaEvent_t)p.ev_prep, sd);
    }
#endif
    tail_forward(c, road, vehicle, navigation, out512, sd, 1);       // feature MLPs + their GRUs, next to the image tower
#ifndef CDRA_EMU
    if (p.v2.on) {               // bf16 perf mode: legacy stem + pool, then the v2 tower (padded planes, TMA tiles)
        if (p.v2.stem_on) v2::stem_forward(c, (const uint8_t*)image);
        else if (u8) tower_forward<bf16, uint8_t>(c, (const uint8_t*)image, true);
        else tower_forward<bf16, float>(c, (const float*)image, true);
        if (sd != c.stream) cudaStreamWaitEvent(c.stream, (cudaEvent_t)p.ev_prep, 0);     // a full dependency: the tower's PDL prologues read the prepared operands
        v2::tower_forward(c);
    } else
#endif
    if (bf && u8) tower_forward<bf16, uint8_t>(c, (const uint8_t*)image);
    else if (bf) tower_forward<bf16, float>(c, (const float*)image);
    else if (u8) tower_forward<float, uint8_t>(c, (const uint8_t*)image);
    else tower_forward<float, float>(c, (const float*)image);
    side_join(c, sd);
    tail_forward(c, road, vehicle, navigation, out512, c.stream, 2);
    return check_launch("dynamics_forward");
}

int cdra_dynamics_backward(cdra_plan_t* plan, const float* params, const void* image, const float* road,
                           const float* vehicle, const float* navigation, const float* d_out512, float* grads,
                           void* workspace, void* stream) {
    if (!plan || !params || !image || !road || !vehicle || !navigation || !d_out512 || !grads || !workspace)
        return fail(CDRA_ERR_BADARG, "null argument");
    const Plan& p = *plan->p;
    RunCtx c{&p, (char*)workspace, params, nullptr, grads, (cudaStream_t)stream, 1};
    t_gemm_tf32 = p.cfg.dtype == CDRA_DTYPE_BF16 && getenv("CDRA_NO_TF32") == nullptr;
    zero_async(grads, (size_t)p.dyn_params.size * 4, c.stream);
    zero_async(c.ws, p.zero_bytes, c.stream);      // BN-backward sums (the forward sums are already folded into aff/bnp)
    tail_backward(c, road, vehicle, navigation, d_out512);
    const bool bf = p.cfg.dtype == CDRA_DTYPE_BF16, u8 = p.cfg.image_u8 != 0;
#ifndef CDRA_EMU
    if (p.v2.on) {
        v2::tower_backward(c);
        if (p.v2.stem_on) v2::stem_backward(c, (const uint8_t*)image);
        else if (u8) tower_backward<bf16, uint8_t>(c, (const uint8_t*)image, true);
        else tower_backward<bf16, float>(c, (const float*)image, true);
    } else
#endif
    if (bf && u8) tower_backward<bf16, uint8_t>(c, (const uint8_t*)image);
    else if (bf) tower_backward<bf16, float>(c, (const float*)image);
    else if (u8) tower_backward<float, uint8_t>(c, (const uint8_t*)image);
    else tower_backward<float, float>(c, (const float*)image);
    side_join(c, side_stream(c));
    return check_launch("dynamics_backward");
}

int cdra_debug_gemm(int ta, int tb, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias,
                    int M, int N, int K, int accumulate, int tensor_core, void* stream) {
    if (!A || !B || !C || M < 1 || N < 1 || K < 1) return fail(CDRA_ERR_BADARG, "bad gemm argument");
    t_gemm_tf32 = tensor_core != 0;
    gemm((cudaStream_t)stream, ta != 0, tb != 0, A, lda, B, ldb, C, ldc, bias, M, N, K, accumulate != 0);
    return check_launch("debug_gemm");
}

int cdra_debug_umma_selftest(const void* X, const void* Y, float* C, int rows, int Mw, int Nw, void* stream) {
    if (!X || !Y || !C || rows < 64 || rows % 64 || (Mw != 128 && Mw != 256) || Nw < 16 || Nw % 16 || Nw > 256 || (Mw / 128) * Nw > 512)
        return fail(CDRA_ERR_BADARG, "bad umma selftest shape");
#ifndef CDRA_EMU
    v2::UmmaTestArgs a{(const bf16*)X, (const bf16*)Y, C, rows, Mw, Nw};
    const int smem = (Mw / 64 + (Nw + 63) / 64) * 64 * 128 + 1024;
    cudaFuncSetAttribute(v2::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    CDRA_LAUNCH(v2::umma_selftest_kernel, dim3(1), dim3(256), smem, (cudaStream_t)stream, a);
    return check_launch("debug_umma_selftest");
#else
    return fail(CDRA_ERR_BADARG, "no tensor cores in the CPU logic-check build");
#endif
}

int cdra_debug_umma_selftest_k(const void* A, const void* B, float* C, int Mw, int Nw, int Kw, void* stream) {
    if (!A || !B || !C || (Mw != 128 && Mw != 256) || Nw < 16 || Nw % 16 || Nw > 256 || (Mw / 128) * Nw > 512 || Kw < 64 || Kw % 64 || Kw > 256)
        return fail(CDRA_ERR_BADARG, "bad umma selftest shape");
#ifndef CDRA_EMU
    v2::UmmaTestKArgs a{(const bf16*)A, (const bf16*)B, C, Mw, Nw, Kw};
    const int smem = (Mw + Nw) * (Kw / 64) * 128 + 1024;
    cudaFuncSetAttribute(v2::umma_selftest_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    CDRA_LAUNCH(v2::umma_selftest_k_kernel, dim3(1), dim3(256), smem, (cudaStream_t)stream, a);
    return check_launch("debug_umma_selftest_k");
#else
    return fail(CDRA_ERR_BADARG, "no tensor cores in the CPU logic-check build");
#endif
}

int cdra_debug_set(const char* key, int value) {
    if (!key) return fail(CDRA_ERR_BADARG, "null key");
#ifndef CDRA_EMU
    if (std::string(key) == "tc") { v2::tc_override() = value; return CDRA_OK; }
    if (std::string(key) == "fwd_tc") { v2::fwd_tc_override() = value; return CDRA_OK; }
    if (std::string(key) == "fused") { v2::fused_override() = value; return CDRA_OK; }
    if (std::string(key) == "pwg") { v2::pwg_override() = value; return CDRA_OK; }
    if (std::string(key) == "dw_band") { v2::dw_band_cap() = value < 0 ? 0 : value; return CDRA_OK; }
#endif
    return fail(CDRA_ERR_BADARG, "unknown debug key");
}

int cdra_debug_stem_backward(cdra_plan_t* plan, const float* params, const void* image, float* grads, void* workspace,
                             int legacy, void* stream) {
    if (!plan || !params || !image || !grads || !workspace) return fail(CDRA_ERR_BADARG, "null argument");
    const Plan& p = *plan->p;
    RunCtx c{&p, (char*)workspace, params, nullptr, grads, (cudaStream_t)stream, 1};
    zero_async(grads, (size_t)p.dyn_params.size * 4, c.stream);
    zero_async(c.ws, p.zero_bytes, c.stream);
    const bool bf = p.cfg.dtype == CDRA_DTYPE_BF16, u8 = p.cfg.image_u8 != 0;
#ifndef CDRA_EMU
    if (!legacy) {
        if (!p.v2.stem_on) return fail(CDRA_ERR_BADARG, "tensor-core stem not active for this plan");
        v2::stem_backward(c, (const uint8_t*)image);
        return check_launch("debug_stem_backward");
    }
#endif
    if (bf && u8) tower_backward<bf16, uint8_t>(c, (const uint8_t*)image, true);
    else if (bf) tower_backward<bf16, float>(c, (const float*)image, true);
    else if (u8) tower_backward<float, uint8_t>(c, (const uint8_t*)image, true);
    else tower_backward<float, float>(c, (const float*)image, true);
    return check_launch("debug_stem_backward");
}

int cdra_policy_head_loss_fwd_bwd(cdra_plan_t* plan, const float* params, float* state, const float* x512,
                                  const float* actions_eval, const float* actions_jac, const float* logp_old, const float* adv,
                                  const float* true_speed, const float* true_sim, float clip_ratio, float ent_coef,
                                  int training, float grad_scale, float* scalars_out, float* head_out, float* d_x512,
                                  float* grads, void* workspace, void* stream) {
    if (!plan || !params || !x512 || !actions_eval || !logp_old || !adv || !true_speed || !true_sim || !scalars_out ||
        !head_out || !workspace) return fail(CDRA_ERR_BADARG, "null argument");
    if (grads && !d_x512) return fail(CDRA_ERR_BADARG, "d_x512 required with grads");
    if (!training && !state) return fail(CDRA_ERR_BADARG, "inference needs the moving statistics");
    return run_head(true, plan, params, state, x512, actions_eval, actions_jac, logp_old, adv, nullptr, true_speed, true_sim,
                    clip_ratio, ent_coef, training, grad_scale, scalars_out, head_out, d_x512, grads, (char*)workspace,
                    (cudaStream_t)stream);
}

int cdra_value_head_loss_fwd_bwd(cdra_plan_t* plan, const float* params, float* state, const float* x512,
                                 const float* returns_be, const float* true_speed, const float* true_sim, int training,
                                 float grad_scale, float* scalars_out, float* head_out, float* d_x512, float* grads,
                                 void* workspace, void* stream) {
    if (!plan || !params || !x512 || !returns_be || !true_speed || !true_sim || !scalars_out || !head_out || !workspace)
        return fail(CDRA_ERR_BADARG, "null argument");
    if (grads && !d_x512) return fail(CDRA_ERR_BADARG, "d_x512 required with grads");
    if (!training && !state) return fail(CDRA_ERR_BADARG, "inference needs the moving statistics");
    return run_head(false, plan, params, state, x512, nullptr, nullptr, nullptr, nullptr, returns_be, true_speed, true_sim, 0.f,
                    0.f, training, grad_scale, scalars_out, head_out, d_x512, grads, (char*)workspace, (cudaStream_t)stream);
}

int cdra_gae(const float* rewards, const float* values_be, const float* last_value_be, double gamma, double lambda_,
             float scale, int bs, int T, float* returns_be_out, float* adv_out, void* stream) {
    if (!rewards || !values_be || !last_value_be || !returns_be_out || !adv_out) return fail(CDRA_ERR_BADARG, "null argument");
    if (bs < 1 || T < 1 || T > 8192) return fail(CDRA_ERR_SHAPE, "bad bs / T");
    GaeArgs a{rewards, values_be, last_value_be, gamma, gamma * lambda_, (float)gamma, scale, bs, T, returns_be_out, adv_out};
    const size_t smem = (size_t)(4 * (T + 1)) * sizeof(float);
#ifndef CDRA_EMU
    if (smem > 48 * 1024) cudaFuncSetAttribute(gae_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    CDRA_LAUNCH(gae_kernel, dim3(bs), dim3(32), smem, (cudaStream_t)stream, a);
    return check_launch("gae");
}

int cdra_clip_adam(float* params, const float* grads, float* m, float* v, const int64_t* tensor_offsets, int n_tensors,
                   int64_t total, float clip_norm, float lr, float beta1, float beta2, float eps, int64_t step,
                   float grad_scale, float* norms_out, void* stream) {
    if (!params || !grads || !m || !v || total < 1 || step < 1) return fail(CDRA_ERR_BADARG, "bad argument");
    if (clip_norm > 0.f && (!tensor_offsets || !norms_out || n_tensors < 1)) return fail(CDRA_ERR_BADARG, "clipping needs offsets + norms");
    AdamArgs a; memset(&a, 0, sizeof a);
    a.p = params; a.g = grads; a.m = m; a.v = v; a.offs = tensor_offsets; a.n_tensors = n_tensors;
    a.clip = clip_norm; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_scale = grad_scale; a.norms = norms_out; a.total = total;
    // Keras Adam: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    a.lr_t = (float)((double)lr * std::sqrt(1.0 - std::pow((double)beta2, (double)step)) / (1.0 - std::pow((double)beta1, (double)step)));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = cdiv(total, kAdamChunk);
    if (clip_norm > 0.f) {
        zero_async(norms_out, (size_t)n_tensors * 4, st);
        CDRA_LAUNCH(sqnorm_kernel, dim3(blocks), dim3(256), 0, st, a);
    }
    CDRA_LAUNCH(adam_kernel, dim3(blocks), dim3(256), 0, st, a);
    return check_launch("clip_adam");
}

int cdra_grad_norms(const float* grads, const int64_t* tensor_offsets, int n_tensors, int64_t total, float grad_scale,
                    float* sq_norms_out, void* stream) {
    if (!grads || !tensor_offsets || !sq_norms_out || n_tensors < 1 || total < 1) return fail(CDRA_ERR_BADARG, "bad argument");
    AdamArgs a; memset(&a, 0, sizeof a);
    a.g = grads; a.offs = tensor_offsets; a.n_tensors = n_tensors; a.grad_scale = grad_scale; a.norms = sq_norms_out; a.total = total;
    cudaStream_t st = (cudaStream_t)stream;
    zero_async(sq_norms_out, (size_t)n_tensors * 4, st);
    CDRA_LAUNCH(sqnorm_kernel, dim3((unsigned)cdiv(total, kAdamChunk)), dim3(256), 0, st, a);
    return check_launch("grad_norms");
}

}  // extern "C"

// ----------------------------------------------------------------------------------------------- NCCL (resolved at run time)
#ifndef CDRA_EMU
#include <dlfcn.h>
struct NcclId { char internal[128]; };        // == ncclUniqueId
namespace {
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value */ NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
}
static NcclApi* nccl_api(std::string& err) {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.handle) break; }   // the copy torch loaded
        for (const char* n : names) { if (api.handle) break; api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
        if (api.handle) {
            api.GetUniqueId = (int (*)(void*))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(api.handle, "ncclCommInitRank");
            api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
            api.CommDestroy = (int (*)(void*))dlsym(api.handle, "ncclCommDestroy");
            api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) { err = "libnccl.so.2 not found / incomplete"; return nullptr; }
    return &api;
}
#endif
struct cdra_comm { void* comm; int world, rank; };

extern "C" {
int cdra_comm_unique_id(void* id_out) {
    if (!id_out) return fail(CDRA_ERR_BADARG, "null argument");
#ifndef CDRA_EMU
    std::string err; NcclApi* api = nccl_api(err);
    if (!api) return fail(CDRA_ERR_NCCL, err);
    const int rc = api->GetUniqueId(id_out);
    if (rc != 0) return fail(CDRA_ERR_NCCL, std::string("ncclGetUniqueId: ") + (api->GetErrorString ? api->GetErrorString(rc) : "?"));
    return CDRA_OK;
#else
    return fail(CDRA_ERR_NCCL, "no NCCL in the CPU logic-check build");
#endif
}
int cdra_comm_create(const void* id, int world, int rank, cdra_comm_t** out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(CDRA_ERR_BADARG, "bad argument");
#ifndef CDRA_EMU
    std::string err; NcclApi* api = nccl_api(err);
    if (!api) return fail(CDRA_ERR_NCCL, err);
    NcclId nid; memcpy(&nid, id, sizeof nid);
    void* comm = nullptr;
    const int rc = api->CommInitRank(&comm, world, nid, rank);
    if (rc != 0) return fail(CDRA_ERR_NCCL, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(rc) : "?"));
    *out = new cdra_comm{comm, world, rank};
    return CDRA_OK;
#else
    return fail(CDRA_ERR_NCCL, "no NCCL in the CPU logic-check build");
#endif
}
void cdra_comm_destroy(cdra_comm_t* c) {
#ifndef CDRA_EMU
    if (c) { std::string err; NcclApi* api = nccl_api(err); if (api && c->comm) api->CommDestroy(c->comm); delete c; }
#else
    delete c;
#endif
}
int cdra_allreduce_grads(cdra_comm_t* c, float* grads, int64_t count, void* stream) {
    if (!c || !grads || count < 1) return fail(CDRA_ERR_BADARG, "bad argument");
#ifndef CDRA_EMU
    std::string err; NcclApi* api = nccl_api(err);
    if (!api) return fail(CDRA_ERR_NCCL, err);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const int rc = api->AllReduce(grads, grads, (size_t)count, /* ncclFloat32 */ 7, /* ncclSum */ 0, c->comm, (cudaStream_t)stream);
    if (rc != 0) return fail(CDRA_ERR_NCCL, std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(rc) : "?"));
    return CDRA_OK;
#else
    return fail(CDRA_ERR_NCCL, "no NCCL in the CPU logic-check build");
#endif
}
}  // extern "C"

extern "C" {
int64_t cdra_launch_count(void) { return (int64_t)g_launches.load(); }
void cdra_profile_enable(int on) { g_prof = on != 0; }
void cdra_profile_reset(void) { g_entries.clear(); }
int cdra_profile_report(char* buf, int cap) {
    std::string s;
#ifndef CDRA_EMU
    for (auto& kv : g_entries) {
        const char* name = nullptr;
        if (cudaFuncGetName(&name, kv.first) != cudaSuccess || !name) name = "?";
        char line[640];
        snprintf(line, sizeof line, "%s\t%lld\t%.6f\t%.0f\n", name, kv.second.count, kv.second.ms, kv.second.bytes);
        s += line;
    }
#endif
    if (buf && cap > 0) { strncpy(buf, s.c_str(), cap - 1); buf[cap - 1] = 0; }
    return (int)s.size();
}

int cdra_gather_rows(const void* src, const int64_t* index, int64_t n, int64_t row_bytes, void* dst, void* stream) {
    if (!src || !index || !dst || n < 1 || row_bytes < 1) return fail(CDRA_ERR_BADARG, "bad argument");
    GatherArgs a{(const char*)src, index, n, row_bytes, (char*)dst};
    unsigned gy = cdiv(row_bytes, 256 * 16 * 8); if (gy < 1) gy = 1; if (gy > 64) gy = 64;
    CDRA_LAUNCH(gather_rows_kernel, dim3((unsigned)n, gy), dim3(256), 0, (cudaStream_t)stream, a);
    return check_launch("gather_rows");
}

int cdra_gather_rows_multi(const void* const* srcs, void* const* dsts, const int64_t* row_bytes, int n_tensors, const int64_t* index,
                           int64_t n, void* stream) {
    if (!srcs || !dsts || !row_bytes || !index || n < 1 || n_tensors < 1 || n_tensors > kGatherMax) return fail(CDRA_ERR_BADARG, "bad argument");
    GatherMultiArgs a; memset(&a, 0, sizeof a);
    a.index = index; a.n = n; a.nt = n_tensors;
    int y = 0;
    for (int k = 0; k < n_tensors; ++k) {
        if (!srcs[k] || !dsts[k] || row_bytes[k] < 1) return fail(CDRA_ERR_BADARG, "bad argument");
        a.src[k] = (const char*)srcs[k]; a.dst[k] = (char*)dsts[k]; a.row_bytes[k] = row_bytes[k];
        int gy = (int)cdiv(row_bytes[k], 256 * 16 * 8); if (gy < 1) gy = 1; if (gy > 64) gy = 64;
        a.y0[k] = y; y += gy;
    }
    a.y0[n_tensors] = y;
    CDRA_LAUNCH(gather_rows_multi_kernel, dim3((unsigned)n, (unsigned)y), dim3(256), 0, (cudaStream_t)stream, a);
    return check_launch("gather_rows_multi");
}

int cdra_debug_timeline(uint64_t* out64) {
    uint64_t* out32 = out64;
    if (!out32) return fail(CDRA_ERR_BADARG, "null argument");
#ifndef CDRA_EMU
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out32, v2::g_pwg_ts, 16 * 8);
    cudaMemcpyFromSymbol(out32 + 16, v2::g_bf_ts, 16 * 8);
    cudaMemcpyFromSymbol(out32 + 32, v2::g_tc_ts, 16 * 8);
    cudaMemcpyFromSymbol(out32 + 48, v2::g_bf_ts, 16 * 8, 16 * 8);
    return check_launch("debug_timeline");
#else
    for (int i = 0; i < 64; ++i) out32[i] = 0;
    return CDRA_OK;
#endif
}

int cdra_augment(const void* image, int image_u8, int64_t frames, int height, int width, const cdra_augment_params* params,
                 const uint8_t* dropout_mask, float* out, void* scratch, void* stream) {
#ifdef CDRA_EMU
    (void)image; (void)image_u8; (void)frames; (void)height; (void)width; (void)params; (void)dropout_mask; (void)out; (void)scratch; (void)stream;
    return fail(CDRA_ERR_BADARG, "cdra_augment needs the CUDA build");
#else
    if (!image || !params || !out || !scratch || frames < 1 || height < 1 || width < 1) return fail(CDRA_ERR_BADARG, "bad argument");
    const cdra_augment_params& p = *params;
    if (p.blur_size != 0 && p.blur_size != 3 && p.blur_size != 5) return fail(CDRA_ERR_BADARG, "blur_size must be 0, 3 or 5");
    if (p.normalize && p.group < 1) return fail(CDRA_ERR_BADARG, "group must be >= 1");
    if (p.dropout_size > 0 && !dropout_mask) return fail(CDRA_ERR_BADARG, "dropout_mask missing");
    if ((long long)height * width >= (1 << 27)) return fail(CDRA_ERR_BADARG, "frame too large for the per-pixel hash");
    cudaStream_t st = (cudaStream_t)stream;
    aug::AugArgs a; memset(&a, 0, sizeof a);
    a.img = image; a.u8 = image_u8; a.frames = frames; a.H = height; a.W = width; a.p = p; a.dropout_mask = dropout_mask; a.out = out;
    const long long groups = p.normalize ? (frames + p.group - 1) / p.group : 1;
    a.mean = (float*)scratch; a.kmin = (uint32_t*)scratch + 3 * frames; a.kmax = a.kmin + groups;
    if (p.normalize) { cudaMemsetAsync(a.kmin, 0xff, (size_t)groups * 4, st); cudaMemsetAsync(a.kmax, 0, (size_t)groups * 4, st); }
    const dim3 grid((unsigned)cdiv((long long)height * width, 256), (unsigned)frames);
    if (p.jitter) CDRA_LAUNCH(aug::aug_mean_kernel, dim3((unsigned)frames), dim3(256), 0, st, a);
    CDRA_LAUNCH(aug::aug_main_kernel, grid, dim3(256), 0, st, a);
    if (p.normalize || p.cutout_size > 0 || p.dropout_size > 0) CDRA_LAUNCH(aug::aug_finish_kernel, grid, dim3(256), 0, st, a);
    return check_launch("augment");
#endif
}

}  // extern "C"
