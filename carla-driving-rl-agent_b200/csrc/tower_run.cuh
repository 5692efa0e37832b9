// Host-side launch sequences for the image tower (forward and backward).
#pragma once
#include "plan.h"
#include <type_traits>
#include <cstdlib>
#include "tower_fwd.cuh"
#include "tower_bwd.cuh"
#include "tower_opt.cuh"
#include "pw_mma.cuh"

namespace cdra {

struct RunCtx {
    const Plan* p;
    char* ws;
    const float* params;     // dynamics trainable arena
    float* state;            // dynamics state arena (moving stats); may be null in backward
    float* grads;            // dynamics gradient arena (backward)
    cudaStream_t stream;
    int training;
};

// debugging aid: CDRA_NO_MMA=1 routes bf16 pointwise convs through the CUDA-core kernels (A/B parity checks)
inline bool use_mma() { static const bool v = getenv("CDRA_NO_MMA") == nullptr; return v; }

inline unsigned* counter_ptr(const RunCtx& c, int idx) { return (unsigned*)(c.ws + c.p->counters_off) + idx; }

inline BnTables tables_of(const RunCtx& c, const WsTensor& t) {
    BnTables tb;
    tb.fst = (double2*)(c.ws + t.fst); tb.bst = (double2*)(c.ws + t.bst);
    tb.aff = (float2*)(c.ws + t.aff); tb.bnp = (float2*)(c.ws + t.bnp);
    return tb;
}
inline BnLayer bn_of(const RunCtx& c, const BnConv& l, int unbiased = 1) {
    BnLayer b;
    b.gamma = c.params + l.g; b.beta = c.params + l.be;
    b.mov_mean = c.state ? c.state + l.mm : nullptr; b.mov_var = c.state ? c.state + l.mv : nullptr;
    b.counter = nullptr;          // BN finalisation runs as its own launch (launch_bn_finalize), no in-kernel ticket
    b.training = c.training; b.unbiased = unbiased;
    return b;
}
inline void launch_bn_finalize(const RunCtx& c, const BnConv& l, const WsTensor& dst, const ColMap& cm, double n) {
    if (!c.training) return;
    BnFinArgs a;
    a.cm = cm; a.ld = dst.C; a.n = n;
    a.tb.fst = (double2*)(c.ws + dst.fst); a.tb.bst = (double2*)(c.ws + dst.bst);
    a.tb.aff = (float2*)(c.ws + dst.aff); a.tb.bnp = (float2*)(c.ws + dst.bnp);
    a.bn.gamma = c.params + l.g; a.bn.beta = c.params + l.be;
    a.bn.mov_mean = c.state ? c.state + l.mm : nullptr; a.bn.mov_var = c.state ? c.state + l.mv : nullptr;
    a.bn.counter = nullptr; a.bn.training = c.training; a.bn.unbiased = 1;
    CDRA_LAUNCH(bn_finalize_kernel, dim3((cm.n + 255) / 256), dim3(256), 0, c.stream, a);
}
inline ActView view_of(const RunCtx& c, const WsTensor& t, int coff, bool clamp) {
    ActView v; v.data = c.ws + t.data; v.ld = t.C; v.coff = coff;
    v.aff = t.tables ? (const float2*)(c.ws + t.aff) : nullptr; v.clamp = clamp ? 1 : 0;
    return v;
}
inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

template <typename T>
void launch_pw_fwd(const RunCtx& c, const BnConv& l, const ActView& in, int Rt, const WsTensor& dst, const ColMap& cm) {
    PwArgs<T> a;
    a.in = in; a.K = l.K; a.Rt = Rt; a.w = c.params + l.w; a.bias = c.params + l.b; a.cm = cm;
    a.out = (T*)(c.ws + dst.data); a.ldo = dst.C; a.tb = tables_of(c, dst); a.bn = bn_of(c, l); a.do_stats = c.training ? 1 : 0;
    prof_bytes(4.0 * Rt * (l.K + cm.n) * sizeof(T));          // read input once, write raw output once
#ifndef CDRA_EMU
    if constexpr (std::is_same<T, bf16>::value) if (use_mma()) {               // tensor-core path (pw_mma.cuh)
        PwMmaFwdArgs pa; pa.a = a; pa.wt = (const bf16*)(c.ws + l.wt); pa.Kp = l.Kp;
        // 64-column tiles: 80 registers / 35 KB smem -> 3 CTAs per SM (the 128-wide variant is register-bound at 2);
        // wide layers re-read their A tile from L2 once per 64 output columns
        auto k64 = pw_fwd_mma_kernel<64>;
        static bool carve = (cudaFuncSetAttribute(k64, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true);
        (void)carve;
        CDRA_LAUNCH(k64, dim3(cdiv(Rt, kMmTM), kT, cdiv(cm.n, 64)), dim3(256), 0, c.stream, pa);
        launch_bn_finalize(c, l, dst, cm, (double)Rt);
        return;
    }
#endif
    dim3 grid(cdiv(Rt, kPwTM), kT, cdiv(cm.n, kPwTN));
    auto k = pw_fwd_kernel<T>;
    CDRA_LAUNCH(k, grid, dim3(256), 0, c.stream, a);
    launch_bn_finalize(c, l, dst, cm, (double)Rt);
}

#ifndef CDRA_EMU
inline void launch_wprep(const RunCtx& c) {
    const Plan& p = *c.p;
    WPrepArgs a; a.n = 0;
    auto add = [&](const BnConv& l) {
        WPrepLayer& L = a.l[a.n++];
        L.w = c.params + l.w; L.wt = (bf16*)(c.ws + l.wt); L.wn = (bf16*)(c.ws + l.wn);
        L.K = l.K; L.N = l.N; L.Kp = l.Kp; L.Np = l.Np; L.split = l.split;
    };
    for (const Unit& u : p.units) { add(u.pw1); add(u.pw2); if (u.stride == 2) add(u.scpw); }
    add(p.head);
    CDRA_LAUNCH(wprep_kernel, dim3(a.n, 8), dim3(256), 0, c.stream, a);
}
#endif

template <typename T>
void launch_dw_fwd(const RunCtx& c, const BnConv& l, const ActView& in, const Unit& u, int C, const WsTensor& dst) {
    DwArgs<T> a;
    a.in = in; a.B = c.p->B; a.Hi = u.Hi; a.Wi = u.Wi; a.Ho = u.Ho; a.Wo = u.Wo; a.C = C; a.stride = u.stride;
    a.pad_t = u.pad_t; a.pad_l = u.pad_l; a.w = c.params + l.w; a.bias = c.params + l.b;
    a.out = (T*)(c.ws + dst.data); a.tb = tables_of(c, dst); a.bn = bn_of(c, l);
    int lanes_c = ((C / 2) + 31) & ~31; if (lanes_c > 256) lanes_c = 256;
    a.ppb = 16 * (256 / lanes_c);
    prof_bytes(4.0 * a.B * ((double)u.Hi * u.Wi + (double)u.Ho * u.Wo) * C * sizeof(T));
    dim3 grid(cdiv((long long)a.B * u.Ho, 256 / lanes_c), kT);      // row-sweep kernels: one output row per thread
    if (u.stride == 1) { auto k = dw_fwd_row_kernel<T, 1>; CDRA_LAUNCH(k, grid, dim3(256), 0, c.stream, a); }
    else { auto k = dw_fwd_row_kernel<T, 2>; CDRA_LAUNCH(k, grid, dim3(256), 0, c.stream, a); }
    launch_bn_finalize(c, l, dst, ColMap{C, 0, 0, 0}, (double)a.B * u.Ho * u.Wo);
}

template <typename T, typename TIn>
void tower_forward(const RunCtx& c, const TIn* image, bool stem_only = false) {
    const Plan& p = *c.p;
    const int B = p.B;
#ifndef CDRA_EMU
    if constexpr (std::is_same<T, bf16>::value) if (!stem_only) launch_wprep(c);
#endif
    {   // stem conv (+BN statistics)                                  core/architectures.py:159-160
        const WsTensor& ts = p.tensors[p.t_stem];
        StemArgs<T, TIn> a;
        a.img = image; a.B = B; a.H = p.H; a.W = p.W; a.Ho = p.Hs; a.Wo = p.Ws;
        a.w = c.params + p.stem.w; a.bias = c.params + p.stem.b; a.out = (T*)(c.ws + ts.data);
        a.tb = tables_of(c, ts); a.bn = bn_of(c, p.stem);
        dim3 grid(cdiv(ts.Rt, 128 * kStemPPT), kT);
        prof_bytes(4.0 * B * ((double)p.H * p.W * 3 * sizeof(TIn) + (double)p.Hs * p.Ws * kStemC * sizeof(T)));
        auto k = stem_fwd_kernel<T, TIn>;
        CDRA_LAUNCH(k, grid, dim3(128), 0, c.stream, a);
        launch_bn_finalize(c, p.stem, ts, ColMap{kStemC, 0, 0, 0}, (double)ts.Rt);
    }
    {   // BN + ReLU6 on load, maxpool 3x3 s2 SAME                      :160-161
        const WsTensor& ts = p.tensors[p.t_stem];
        const WsTensor& tp = p.tensors[p.t_pool];
        PoolArgs<T> a;
        a.in = view_of(c, ts, 0, true); a.B = B; a.Hi = p.Hs; a.Wi = p.Ws; a.Ho = p.Hp; a.Wo = p.Wp; a.C = kStemC;
        a.pad_t = p.pool_pad_t; a.pad_l = p.pool_pad_l; a.out = (T*)(c.ws + tp.data);
        dim3 grid(cdiv((long long)tp.Rt * (kStemC / 2), 256), kT);
        auto k = pool_fwd_kernel<T>;
        CDRA_LAUNCH(k, grid, dim3(256), 0, c.stream, a);
    }
    if (stem_only) return;           // v2 tower (bf16): the units / head run through v2_run.cuh
    for (const Unit& u : p.units) {                                     // :120-151
        const WsTensor& tin = p.tensors[u.t_in];
        const WsTensor& r1 = p.tensors[u.t_r1];
        const WsTensor& r2 = p.tensors[u.t_r2];
        const WsTensor& out = p.tensors[u.t_out];
        const bool in_clamp = tin.tables;        // unit outputs are BN+ReLU6 outputs; the pool output is plain
        const int sc = u.stride == 2 ? u.cin : u.cin / 2;
        // branch: pw1 -> BN/ReLU6 -> dw -> BN -> pw2 -> BN/ReLU6
        ActView x = view_of(c, tin, u.stride == 2 ? 0 : u.cin / 2, in_clamp);
        launch_pw_fwd<T>(c, u.pw1, x, tin.Rt, r1, ColMap{u.half, 0, 0, 0});
        launch_dw_fwd<T>(c, u.dw, view_of(c, r1, 0, true), u, u.half, r2);
        launch_pw_fwd<T>(c, u.pw2, view_of(c, r2, 0, false), r2.Rt, out, ColMap{u.c - sc, 1, sc / 2, u.half});
        if (u.stride == 2) {
            const WsTensor& rs = p.tensors[u.t_rs];
            launch_dw_fwd<T>(c, u.scdw, view_of(c, tin, 0, in_clamp), u, sc, rs);
            launch_pw_fwd<T>(c, u.scpw, view_of(c, rs, 0, false), rs.Rt, out, ColMap{sc, 1, 0, u.half});
        } else {
            PassArgs<T> a;
            a.in = (const T*)(c.ws + tin.data); a.ldi = tin.C; a.out = (T*)(c.ws + out.data); a.ldo = out.C;
            a.half = u.half; a.Rt = out.Rt;
            a.aff_in = tin.tables ? (const float2*)(c.ws + tin.aff) : nullptr; a.aff_out = (float2*)(c.ws + out.aff);
            a.bnp_in = tin.tables ? (const float2*)(c.ws + tin.bnp) : nullptr; a.bnp_out = (float2*)(c.ws + out.bnp);
            dim3 grid(cdiv((long long)out.Rt * (u.half / 2), 256), kT);
            auto k = pass_fwd_kernel<T>;
            CDRA_LAUNCH(k, grid, dim3(256), 0, c.stream, a);
        }
    }
    {   // head conv 464 -> 768, BN, ReLU6, global average pool        :169-172
        const WsTensor& tin = p.tensors[p.units.back().t_out];
        const WsTensor& th = p.tensors[p.t_head];
        launch_pw_fwd<T>(c, p.head, view_of(c, tin, 0, true), tin.Rt, th, ColMap{p.head.N, 0, 0, 0});
        GapArgs<T> a;
        a.in = view_of(c, th, 0, true); a.B = B; a.HW = th.H * th.W; a.C = th.C; a.out = (float*)(c.ws + p.gap);
        dim3 grid(cdiv((long long)B * th.C, 256), kT);
        auto k = gap_fwd_kernel<T>;
        CDRA_LAUNCH(k, grid, dim3(256), 0, c.stream, a);
    }
}

}  // namespace cdra
