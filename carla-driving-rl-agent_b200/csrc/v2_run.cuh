// v2 tower: host-side launch sequences (forward, backward) and the logical-layout export used by the parity taps.
#pragma once
#ifndef CDRA_EMU
#include <mutex>
#include <set>
#include "plan.h"
#include "tower_run.cuh"
#include "v2_pw.cuh"
#include "v2_dw.cuh"
#include "v2_bwd.cuh"
#include "v2_stem.cuh"
#include "v2_umma.cuh"
#include "v2_pw_tc.cuh"
#include "v3_pw_bwd.cuh"
#include "v4_pwg.cuh"

namespace cdra {
namespace v2 {

inline Tables tables_of(const RunCtx& c, const V2Tensor& t) {
    Tables tb;
    tb.fsum = (double2*)(c.ws + t.fsum); tb.bsum = (double2*)(c.ws + t.bsum);
    tb.aff = (float2*)(c.ws + t.aff); tb.bnp = (float2*)(c.ws + t.bnp);
    return tb;
}
inline LayerP layer_of(const RunCtx& c, const BnConv& l) {
    LayerP L;
    L.w = c.params + l.w; L.b = c.params + l.b; L.g = c.params + l.g; L.be = c.params + l.be;
    L.mm = c.state ? c.state + l.mm : nullptr; L.mv = c.state ? c.state + l.mv : nullptr;
    if (c.grads) { L.dw = c.grads + l.w; L.db = c.grads + l.b; L.dg = c.grads + l.g; L.dbe = c.grads + l.be; }
    else { L.dw = L.db = L.dg = L.dbe = nullptr; }
    L.K = l.K; L.N = l.N;
    return L;
}
inline PwSrc src_of(const RunCtx& c, const V2Tensor& t, bool clamp, int kbase, int layer) {
    PwSrc s;
    s.data = (const bf16*)(c.ws + t.data); s.grad = (bf16*)(c.ws + t.grad);
    s.aff = t.has_bn ? (const float2*)(c.ws + t.aff) : nullptr;
    s.bnp = t.has_bn ? (const float2*)(c.ws + t.bnp) : nullptr;
    s.bsum = (double2*)(c.ws + t.bsum);
    s.cp = t.cp; s.clamp = clamp ? 1 : 0; s.map = SlotMap{t.n0, t.n0p, t.n1}; s.kbase = kbase; s.layer = layer;
    s.sum_lo = 0; s.sum_hi = t.has_bn ? t.cp : 0; s.accumulate = 0;
    return s;
}

// ---- GEMM descriptors of every pointwise launch (same objects drive weight prep, forward and backward)
struct UnitDescs { PwDesc pw1, tail; };

inline void fill_pw(PwDesc& d, const RunCtx& c, const V2Pw& g) {
    d.KP = g.KP; d.NPall = g.NPall;
    d.wf = (bf16*)(c.ws + g.wf); d.wb = (bf16*)(c.ws + g.wb); d.bias = (float*)(c.ws + g.bias);
    d.wfs = (bf16*)(c.ws + g.wfs); d.wbs = (bf16*)(c.ws + g.wbs);
    d.cols.nplanes = g.nplanes; d.cols.gwp = g.gwp;
}
inline PwDesc desc_pw1(const RunCtx& c, int ui) {
    const Plan& p = *c.p; const V2Plan& v = p.v2; const V2Unit& u = v.u[ui]; const Unit& un = p.units[ui];
    PwDesc d; memset(&d, 0, sizeof d);
    fill_pw(d, c, u.pw1);
    const bool in_clamp = u.inB >= 0;            // unit outputs are BN+ReLU6 outputs; the pool output is already activated
    if (un.stride == 2) {
        d.nsrc = 0;
        d.src[d.nsrc++] = src_of(c, v.t[u.inA], in_clamp, 0, 0);
        if (u.inB >= 0) d.src[d.nsrc++] = src_of(c, v.t[u.inB], in_clamp, v.t[u.inA].C(), 0);
        for (int i = 0; i < d.nsrc; ++i) d.src[i].accumulate = 1;     // backward: the shortcut depthwise wrote its share first
    } else {
        d.nsrc = 1; d.src[0] = src_of(c, v.t[u.inB], true, 0, 0);
    }
    d.layer[0] = layer_of(c, un.pw1); d.layer[1] = d.layer[0];
    d.cols.seg0p = u.pw1.gwp; d.cols.seg0n = un.half; d.cols.seg1n = 0; d.cols.interleave = 0;
    return d;
}
inline PwDesc desc_tail(const RunCtx& c, int ui) {
    const Plan& p = *c.p; const V2Plan& v = p.v2; const V2Unit& u = v.u[ui]; const Unit& un = p.units[ui];
    PwDesc d; memset(&d, 0, sizeof d);
    fill_pw(d, c, u.tail);
    const V2Tensor& o = v.t[u.outA];
    d.nsrc = 0;
    d.src[d.nsrc++] = src_of(c, v.t[u.r2], false, 0, 0);
    d.layer[0] = layer_of(c, un.pw2); d.layer[1] = d.layer[0];
    d.cols.interleave = 1;
    d.cols.seg0p = o.n0p; d.cols.seg0n = o.n0; d.cols.seg1n = 0;
    if (un.stride == 2) {
        d.src[d.nsrc++] = src_of(c, v.t[u.rsA], false, 0, 1);
        if (u.rsB >= 0) d.src[d.nsrc++] = src_of(c, v.t[u.rsB], false, v.t[u.rsA].C(), 1);
        d.layer[1] = layer_of(c, un.scpw);
        d.cols.seg1n = o.n1;
    }
    return d;
}
inline PwDesc desc_head(const RunCtx& c) {
    const Plan& p = *c.p; const V2Plan& v = p.v2; const V2Unit& u = v.u.back();
    PwDesc d; memset(&d, 0, sizeof d);
    fill_pw(d, c, v.head_pw);
    d.nsrc = 2;
    d.src[0] = src_of(c, v.t[u.outA], true, 0, 0);
    d.src[1] = src_of(c, v.t[u.outB], true, v.t[u.outA].C(), 0);
    d.layer[0] = layer_of(c, p.head); d.layer[1] = d.layer[0];
    d.cols.seg0p = v.head_pw.gwp; d.cols.seg0n = p.head.N; d.cols.seg1n = 0; d.cols.interleave = 0;
    return d;
}

// the forward (no gradient pointers) and the backward descriptor sets live side by side in the workspace, so that neither
// call invalidates the other's device copy
inline size_t desc_slot(const Plan& p, int i, int which) { return p.v2.desc_off + (size_t)which * 32768 + (size_t)i * sizeof(PwDesc); }
// launch index: 2*ui = pw1, 2*ui+1 = tail, 2*nunits = head
inline const PwDesc* desc_dev(const RunCtx& c, int i) { return (const PwDesc*)(c.ws + desc_slot(*c.p, i, c.grads ? 1 : 0)); }

inline void upload_descs(const RunCtx& c) {
    const Plan& p = *c.p; const int nu = (int)p.v2.u.size(), n = 2 * nu + 1, which = c.grads ? 1 : 0;
    static_assert(3 * kPrepMax * sizeof(PwDesc) <= 64 * 1024 && kPrepMax * sizeof(PwDesc) <= 32768, "descriptor buffers");
    PwDesc* host = (PwDesc*)p.v2.host_descs;
    PwDesc* prev = host + (1 + which) * kPrepMax;    // what the device copy of this set holds
    for (int ui = 0; ui < nu; ++ui) { host[2 * ui] = desc_pw1(c, ui); host[2 * ui + 1] = desc_tail(c, ui); }
    host[2 * nu] = desc_head(c);
    // the descriptors only hold pointers into the caller's arenas / workspace: with the same buffers as in the previous call
    // of the same kind (every SGD step after the first) the device copy is already right
    if (p.v2.desc_ws[which] == c.ws && memcmp(prev, host, (size_t)n * sizeof(PwDesc)) == 0) return;
    memcpy(prev, host, (size_t)n * sizeof(PwDesc));
    p.v2.desc_ws[which] = c.ws;
    cudaMemcpyAsync(c.ws + desc_slot(p, 0, which), prev, (size_t)n * sizeof(PwDesc), cudaMemcpyHostToDevice, c.stream);
}

constexpr int kMaxDynSmem = 226 * 1024;      // 227 KB opt-in limit minus the kernels' few static bytes
inline int& tc_override() { static int v = -1; return v; }      // cdra_debug_set("tc", 0 | 1): A/B parity runs in one process
inline bool use_tc() { static const bool env = getenv("CDRA_NO_TC") == nullptr; return tc_override() < 0 ? env : tc_override() != 0; }
inline int& fwd_tc_override() { static int v = -1; return v; }  // cdra_debug_set("fwd_tc", 0 | 1)
inline bool use_fwd_tc() { static const bool env = getenv("CDRA_NO_FWD_TC") == nullptr; return fwd_tc_override() < 0 ? env : fwd_tc_override() != 0; }
inline int& fused_override() { static int v = -1; return v; }   // cdra_debug_set("fused", 0 | 1)
inline bool use_fused() { static const bool env = getenv("CDRA_NO_FUSED") == nullptr; return fused_override() < 0 ? env : fused_override() != 0; }
inline int& pwg_override() { static int v = -1; return v; }     // cdra_debug_set("pwg", 0 | 1)
inline bool use_pwg() { static const bool env = getenv("CDRA_NO_PWG") == nullptr; return pwg_override() < 0 ? env : pwg_override() != 0; }
// the GEMM family of v4_pwg.cuh takes every pointwise launch whose reduction length reaches this (stage 3 and the head)
inline int pwg_min_k() { static const int v = getenv("CDRA_PWG_MINK") ? atoi(getenv("CDRA_PWG_MINK")) : 192; return v; }
inline int use_timeline() { static const int v = getenv("CDRA_TIMELINE") != nullptr; return v; }
inline int& dw_band_cap() { static int v = 0; return v; }        // cdra_debug_set("dw_band", rows): 0 = automatic
inline bool use_pwg_wgrad() { static const bool env = getenv("CDRA_NO_PWG_WGRAD") == nullptr; return env; }
inline int pwg_min_k_bwd() { static const int v = getenv("CDRA_PWG_MINK_BWD") ? atoi(getenv("CDRA_PWG_MINK_BWD")) : 192; return v; }
// opt a kernel into the full dynamic shared memory ONCE (the attribute call costs microseconds; the launchers run ~500 times a step)
template <typename K>
inline void max_smem_once(K k) {
    static std::mutex mu;
    static std::set<const void*> done;
    std::lock_guard<std::mutex> lock(mu);
    if (done.insert((const void*)k).second) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
}
inline int num_sms() {
    static int n = [] { int d = 0, v = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
    return n;
}

// ---- forward GEMM launch: picks the tile configuration by column count / shared-memory footprint
template <int R, int WM, int WN, int MT, int NBW>
inline bool try_pw_fwd(const RunCtx& c, PwFwdArgs& a, const PwDesc& hd, int colmode, int min_ctas_per_sm) {
    constexpr int NT = WN * NBW * 8;
    int src_row_bytes = 0;
    for (int i = 0; i < hd.nsrc; ++i) src_row_bytes += hd.src[i].cp * 2;
    if (a.x1) src_row_bytes += a.x1cp * 2;
    int gy = 1, splanes = hd.cols.nplanes, swidth = a.cpo;
    a.colmode = colmode;
    if (colmode == 0) { if (hd.NPall > NT) return false; a.ntiles_n = 1; }
    else {                                   // N-tiled inside planes (no pass-through copy)
        a.ntiles_n = (hd.cols.gwp + NT - 1) / NT; gy = hd.cols.nplanes * a.ntiles_n; splanes = 1; swidth = NT;
    }
    const PwFwdSmem L = pw_fwd_smem(R, NT, hd.KP, src_row_bytes, splanes, swidth);
    if (L.total > kMaxDynSmem) return false;
    if (min_ctas_per_sm > 1 && (L.total + 1024) * min_ctas_per_sm > 227 * 1024) return false;
    auto k = pw_fwd_kernel<R, WM, WN, MT, NBW>;
    static bool attr_done = (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), true);
    (void)attr_done;
    const int per_sm = std::max(1, std::min(2, (227 * 1024) / (L.total + 1024)));
    const int tps = (a.Rt + R - 1) / R, ntile = kT * tps;
    int gx = std::max(1, num_sms() * per_sm / gy);
    if (gx > ntile) gx = ntile;
    a.tiles_per_cta = (ntile + gx - 1) / gx;
    gx = (ntile + a.tiles_per_cta - 1) / a.tiles_per_cta;
    CDRA_LAUNCH_PDL(k, dim3(gx, gy), dim3(256), L.total, c.stream, a);
    return true;
}

// plain-output layers (pw1 of every unit) on tcgen05: one CTA per SM, TMEM accumulator, weights resident in the swizzled layout
inline bool try_pw_fwd_tc(const RunCtx& c, PwFwdArgs& a, const PwDesc& hd) {
    constexpr int NT = 512;
    if (hd.cols.nplanes > 2 || (a.gwv & 1)) return false;
    int sum = 0;
    for (int i = 0; i < hd.nsrc; ++i) sum += hd.src[i].cp;
    const int x1cp = a.x1 ? a.x1cp : 0, nq = hd.cols.nplanes * (a.cpo >> 3);
    const PwFwdTcSmem L0 = pw_fwd_tc_smem(hd.KP, hd.NPall, a.cpo, sum, 0, NT, hd.cols.nplanes, x1cp);
    if (L0.np > 256 || L0.nkb > 4 || (NT / nq) < 1 || (NT / (sum >> 3)) < 1) return false;
    if (a.x1 && NT / (2 * ((a.ncopy + 1) >> 1)) < 1) return false;
    const int nbuf = std::min(8, (kMaxDynSmem - L0.total) / L0.raw_stride);
    if (nbuf < 2) return false;
    const PwFwdTcSmem L = pw_fwd_tc_smem(hd.KP, hd.NPall, a.cpo, sum, nbuf, NT, hd.cols.nplanes, x1cp);
    auto k = pw_fwd_tc_kernel<NT>;
    static bool attr_done = (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), true);
    (void)attr_done;
    a.nbuf = nbuf; a.colmode = 0; a.ntiles_n = 1; a.timeline = use_timeline();
    const int tps = (a.Rt + 127) / 128, ntile = kT * tps;
    int gx = std::min(ntile, num_sms());
    a.tiles_per_cta = (ntile + gx - 1) / gx;
    gx = (ntile + a.tiles_per_cta - 1) / a.tiles_per_cta;
    CDRA_LAUNCH_PDL(k, dim3(gx), dim3(NT), L.total, c.stream, a);
    return true;
}

// shared-memory plan of a v4_pwg launch: resident weights when they leave room for >= 3 ring stages, else streamed
inline bool pwg_plan(int kdim, int tab_bytes, int extra, int& ares, int& nstage, int& smem) {
    // bytes in flight hide the load latency (one 16 KB activation block per ring stage): resident weights only when they
    // still leave a deep ring, else the weights stream through the ring next to the activations
    const PwgSmem La = pwg_smem(kdim, tab_bytes, true, 128, 0, extra), Ls = pwg_smem(kdim, tab_bytes, false, 128, 0, extra);
    const int na = std::min(kGMaxStages, (kMaxDynSmem - La.total) / La.stage_bytes);
    const int ns = std::min(kGMaxStages, (kMaxDynSmem - Ls.total) / Ls.stage_bytes);
    if (na >= 5 || na >= ns) { ares = 1; nstage = na; } else { ares = 0; nstage = ns; }
    if (nstage < 3) return false;
    smem = pwg_smem(kdim, tab_bytes, ares != 0, 128, nstage, extra).total;
    return true;
}
inline void pwg_grid(int Rt, int nblk, int& gx, int& tiles_per_cta) {
    const int tps = (Rt + kGRows - 1) / kGRows, ntile = kT * tps;
    const int per_blk = std::max(1, num_sms() / nblk);
    tiles_per_cta = (ntile + per_blk - 1) / per_blk;
    gx = (ntile + tiles_per_cta - 1) / tiles_per_cta;
}

inline bool try_pwg_fwd(const RunCtx& c, PwFwdArgs& a, const PwDesc& hd) {
    if (hd.KP < pwg_min_k() || hd.KP > 768 || (a.gwv & 1)) return false;
    a.nblk = 0;
    for (int p = 0; p < hd.cols.nplanes; ++p)
        for (int s0 = 0; s0 < hd.cols.gwp; s0 += 128) {
            if (a.gwv - s0 <= 0) continue;
            if (a.nblk == kGMaxBlk) return false;
            a.blk[a.nblk++] = GBlock{p * hd.cols.gwp + s0, std::min(128, hd.cols.gwp - s0), p, s0};
        }
    int ares, nstage, smem;
    if (!pwg_plan(hd.KP, ((hd.KP + 63) & ~63) * 8, 0, ares, nstage, smem)) return false;
    a.nbuf = nstage;
    a.timeline = use_timeline();
    int gx; pwg_grid(a.Rt, a.nblk, gx, a.tiles_per_cta);
    // K blocks in flight per producer thread: two ring stages of slack, so that a producer never waits for the MMA round trip
    // (full -> tcgen05.mma -> commit -> empty) of the block it has just handed over
    const int depth = std::max(1, std::min(4, nstage - 2));
    auto launch = [&](auto k) {
        max_smem_once(k);
        CDRA_LAUNCH_PDL(k, dim3(gx, a.nblk), dim3(kGThreads), smem, c.stream, a, ares);
    };
    if (depth == 4) launch(pwg_fwd_kernel<4>); else if (depth == 3) launch(pwg_fwd_kernel<3>);
    else if (depth == 2) launch(pwg_fwd_kernel<2>); else launch(pwg_fwd_kernel<1>);
    if (a.x1) {       // pass-through half of a stride-1 unit
        PassFwdArgs q; q.x1 = a.x1; q.x1cp = a.x1cp; q.x1map = a.x1map; q.out[0] = a.out[0]; q.out[1] = a.out[1]; q.cpo = a.cpo;
        q.ncopy = a.ncopy; q.copy_dst0 = a.copy_dst0; q.rows = (long long)kT * a.Rt;
        prof_bytes(4.0 * a.Rt * 2 * a.ncopy * 2 * 2);
        const long long total = q.rows * q.ncopy;
        CDRA_LAUNCH_PDL(pass_fwd_kernel, dim3((unsigned)std::min<long long>((total + 255) / 256, 8 * num_sms())), dim3(256), 0, c.stream, q);
    }
    return true;
}

inline void launch_pw_fwd(const RunCtx& c, int di, const PwDesc& hd, PwFwdArgs& a, bool allow_full_cols) {
    a.d = desc_dev(c, di);
    a.training = c.training;
    double bytes = 0;
    for (int i = 0; i < hd.nsrc; ++i) bytes += 4.0 * a.Rt * hd.src[i].map.n0 * 2 + 4.0 * a.Rt * hd.src[i].map.n1 * 2;
    for (int pl = 0; pl < hd.cols.nplanes; ++pl) bytes += 4.0 * a.Rt * (hd.cols.seg0n + hd.cols.seg1n + a.ncopy) * 2;
    if (a.x1) bytes += 4.0 * a.Rt * 2 * a.ncopy * 2;
    prof_bytes(bytes);
    bool ok = false;
    if (use_tc() && use_pwg() && try_pwg_fwd(c, a, hd)) return;
    if (allow_full_cols && use_tc() && use_fwd_tc() && try_pw_fwd_tc(c, a, hd)) return;
    if (allow_full_cols) {
        if (hd.NPall <= 64) ok = try_pw_fwd<128, 8, 1, 1, 8>(c, a, hd, 0, 2) || try_pw_fwd<64, 4, 2, 1, 4>(c, a, hd, 0, 1);
        else if (hd.NPall <= 128) ok = try_pw_fwd<64, 4, 2, 1, 8>(c, a, hd, 0, 2) || try_pw_fwd<32, 2, 4, 1, 4>(c, a, hd, 0, 1);
        else if (hd.NPall <= 256) ok = try_pw_fwd<16, 1, 8, 1, 4>(c, a, hd, 0, 1);
    }
    if (!ok && !a.x1) ok = try_pw_fwd<32, 2, 4, 1, 4>(c, a, hd, 1, 1) || try_pw_fwd<32, 2, 4, 1, 2>(c, a, hd, 1, 1);
    if (!ok) fprintf(stderr, "libcdra: no pw_fwd configuration fits (KP=%d NPall=%d)\n", hd.KP, hd.NPall);
}

// output rows per band of a depthwise launch: the whole frame if its double-buffered footprint stays under `limit`, else the
// largest band that does (evened out over the bands)
inline bool dw_pick_band(int cp, const Unit& u, bool backward, int limit, int& band_rows, int& nbands) {
    auto total = [&](int bh) { return dw_smem(cp, u.Hi, u.Wi, u.Ho, u.Wo, 2, backward, bh, u.stride).total; };
    int bh = u.Ho;
    if (total(bh) > limit) {
        bh = 1;
        if (total(bh) > kMaxDynSmem) return false;
        const int lim = total(1) > limit ? kMaxDynSmem : limit;
        while (bh < u.Ho && total(bh + 1) <= lim) ++bh;
    }
    if (dw_band_cap() > 0) bh = std::min(bh, dw_band_cap());     // test switch: force banding on frames that would fit
    nbands = (u.Ho + bh - 1) / bh;
    band_rows = (u.Ho + nbands - 1) / nbands;
    return true;
}

inline void launch_dw_fwd(const RunCtx& c, const BnConv& l, const V2Tensor& in, bool clamp, int kbase, const V2Tensor& out,
                          const Unit& u, int counter) {
    DwArgs a; memset(&a, 0, sizeof a);
    a.in = (const bf16*)(c.ws + in.data); a.aff = in.has_bn ? (const float2*)(c.ws + in.aff) : nullptr; a.clamp = clamp ? 1 : 0;
    a.cp = in.cp; a.map = SlotMap{in.n0, in.n0p, in.n1}; a.kbase = kbase;
    a.B = c.p->B; a.Hi = u.Hi; a.Wi = u.Wi; a.Ho = u.Ho; a.Wo = u.Wo; a.stride = u.stride; a.pad_t = u.pad_t; a.pad_l = u.pad_l;
    a.L = layer_of(c, l);
    a.out = (bf16*)(c.ws + out.data); a.tb = tables_of(c, out);
    a.training = c.training; a.counter = counter_ptr(c, counter);
    typedef void (*KernelT)(const DwArgs);
    KernelT k = nullptr;
    switch (in.cp * 10 + u.stride) {
        case 641: k = dw_fwd_kernel<64, 1>; break;    case 1201: k = dw_fwd_kernel<120, 1>; break;  case 2321: k = dw_fwd_kernel<232, 1>; break;
        case 242: k = dw_fwd_kernel<24, 2>; break;    case 642: k = dw_fwd_kernel<64, 2>; break;
        case 1202: k = dw_fwd_kernel<120, 2>; break;  case 2322: k = dw_fwd_kernel<232, 2>; break;
    }
    if (!k) { fprintf(stderr, "libcdra: no depthwise kernel for cp=%d stride=%d\n", in.cp, u.stride); return; }
    a.nbuf = 2;
    // whole frames when two CTAs of them fit an SM, else the largest row band that does
    if (!dw_pick_band(in.cp, u, false, 110 * 1024, a.band_rows, a.nbands)) { fprintf(stderr, "libcdra: depthwise band does not fit in shared memory (cp=%d)\n", in.cp); return; }
    const DwSmem L = dw_smem(in.cp, u.Hi, u.Wi, u.Ho, u.Wo, a.nbuf, false, a.band_rows, u.stride);
    max_smem_once(k);
    const int per_sm = std::max(1, std::min(2, (227 * 1024) / (L.total + 1024)));
    const int nitems = kT * a.B * a.nbands;
    int gx = std::min(nitems, num_sms() * per_sm);
    a.frames_per_cta = (nitems + gx - 1) / gx;
    gx = (nitems + a.frames_per_cta - 1) / a.frames_per_cta;
    prof_bytes(4.0 * a.B * ((double)u.Hi * u.Wi + (double)u.Ho * u.Wo) * (in.n0 + in.n1) * 2);
    CDRA_LAUNCH_PDL(k, dim3(gx), dim3(kDwThreads), L.total, c.stream, a);
}

// descriptors + bf16 operand matrices of every GEMM; runs BEFORE the stem so that the programmatically launched tower
// kernels (whose prologues read them ahead of pdl_wait) are separated from it by fully ordered launches
inline void tower_prepare(const RunCtx& c) {
    const int nu = (int)c.p->v2.u.size();
    upload_descs(c);
    CDRA_LAUNCH(pw_prep_kernel, dim3(2 * nu + 1, 32), dim3(256), 0, c.stream, desc_dev(c, 0));
}

inline void tower_forward(const RunCtx& c) {
    const Plan& p = *c.p; const V2Plan& v = p.v2;
    const int nu = (int)v.u.size();
    const PwDesc* host = (const PwDesc*)v.host_descs;
    for (int ui = 0; ui < nu; ++ui) {
        const V2Unit& u = v.u[ui]; const Unit& un = p.units[ui];
        const V2Tensor& r1 = v.t[u.r1]; const V2Tensor& r2 = v.t[u.r2];
        const V2Tensor& oA = v.t[u.outA]; const V2Tensor& oB = v.t[u.outB];
        {   // pw1 -> r1
            PwFwdArgs a; memset(&a, 0, sizeof a);
            a.Rt = r1.Rt; a.out[0] = (bf16*)(c.ws + r1.data); a.out[1] = a.out[0]; a.cpo = r1.cp;
            a.tb[0] = tables_of(c, r1); a.tb[1] = a.tb[0]; a.gwv = r1.cp;
            a.counter = counter_ptr(c, u.pw1.counter);
            launch_pw_fwd(c, 2 * ui, host[2 * ui], a, true);
        }
        launch_dw_fwd(c, un.dw, r1, true, 0, r2, un, u.c_dw);
        if (un.stride == 2) {
            const bool in_clamp = u.inB >= 0;
            launch_dw_fwd(c, un.scdw, v.t[u.inA], in_clamp, 0, v.t[u.rsA], un, u.c_scA);
            if (u.inB >= 0) launch_dw_fwd(c, un.scdw, v.t[u.inB], in_clamp, v.t[u.inA].C(), v.t[u.rsB], un, u.c_scB);
        }
        {   // tail: pw2 (+ shortcut pw | pass-through) + shuffle -> planes A, B
            PwFwdArgs a; memset(&a, 0, sizeof a);
            a.Rt = oA.Rt; a.out[0] = (bf16*)(c.ws + oA.data); a.out[1] = (bf16*)(c.ws + oB.data); a.cpo = oA.cp;
            a.tb[0] = tables_of(c, oA); a.tb[1] = tables_of(c, oB);
            a.counter = counter_ptr(c, u.tail.counter);
            if (un.stride == 2) a.gwv = oA.cp;
            else {
                const V2Tensor& x1 = v.t[u.inA];
                a.gwv = oA.n0p;
                a.x1 = (const bf16*)(c.ws + x1.data); a.x1cp = x1.cp; a.x1map = SlotMap{x1.n0, x1.n0p, x1.n1};
                a.x1aff = (const float2*)(c.ws + x1.aff); a.x1bnp = (const float2*)(c.ws + x1.bnp);
                a.ncopy = oA.n1; a.copy_dst0 = oA.n0p;
            }
            launch_pw_fwd(c, 2 * ui + 1, host[2 * ui + 1], a, true);
        }
    }
    {   // head conv + global average pool
        const V2Tensor& th = v.t[v.head];
        PwFwdArgs a; memset(&a, 0, sizeof a);
        a.Rt = th.Rt; a.out[0] = (bf16*)(c.ws + th.data); a.out[1] = a.out[0]; a.cpo = th.cp;
        a.tb[0] = tables_of(c, th); a.tb[1] = a.tb[0]; a.gwv = th.cp;
        a.counter = counter_ptr(c, v.head_pw.counter);
        launch_pw_fwd(c, 2 * nu, host[2 * nu], a, false);
        GapArgs g; memset(&g, 0, sizeof g);
        g.in = (const bf16*)(c.ws + th.data); g.aff = (const float2*)(c.ws + th.aff); g.cp = th.cp; g.C = th.n0; g.HW = th.H * th.W;
        g.F = kT * p.B; g.B = p.B; g.out = (float*)(c.ws + p.gap);
        prof_bytes((double)g.F * g.HW * g.C * 2);
        CDRA_LAUNCH_PDL(gap_fwd_kernel, dim3(cdiv((long long)g.F * (th.cp / 8), 256)), dim3(256), 0, c.stream, g);
    }
}

// ================================================================================================ backward
template <int R, int WM, int WN, int MT, int NBW>
inline bool try_pw_dgrad(const RunCtx& c, PwBwdArgs& a, const PwDesc& hd, int nbuf, int direct, int min_ctas_per_sm) {
    constexpr int KT = WN * NBW * 8, NT = WM * WN * 32;
    int max_cp = 0, gy = 0;
    for (int i = 0; i < kMaxSrc; ++i) a.ntiles_k[i] = 0;
    for (int i = 0; i < hd.nsrc; ++i) { max_cp = std::max(max_cp, hd.src[i].cp); a.ntiles_k[i] = (hd.src[i].cp + KT - 1) / KT; gy += a.ntiles_k[i]; }
    if (direct && a.x1) return false;
    const PwDgradSmem L = pw_dgrad_smem(R, KT, hd.NPall, hd.cols.nplanes, a.cpo, max_cp, a.x1 ? a.x1cp : 0, nbuf, direct, NT);
    if (NT > 256 && min_ctas_per_sm > 1) return false;
    if (L.total > kMaxDynSmem) return false;
    if (min_ctas_per_sm > 1 && (L.total + 1024) * min_ctas_per_sm > 227 * 1024) return false;
    auto k = pw_dgrad_kernel<R, WM, WN, MT, NBW>;
    static bool attr_done = (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), true);
    (void)attr_done;
    a.nbuf = nbuf; a.direct = direct;
    const int per_sm = NT > 256 ? 1 : std::max(1, std::min(2, (227 * 1024) / (L.total + 1024)));
    const int tps = (a.Rt + R - 1) / R, ntile = kT * tps;
    int gx = std::max(1, num_sms() * per_sm / gy);
    if (gx > ntile) gx = ntile;
    a.tiles_per_cta = (ntile + gx - 1) / gx;
    gx = (ntile + a.tiles_per_cta - 1) / a.tiles_per_cta;
    CDRA_LAUNCH_PDL(k, dim3(gx, gy), dim3(NT), L.total, c.stream, a);
    return true;
}

template <int NBW>
inline bool try_pw_wgrad(const RunCtx& c, PwBwdArgs& a, const PwDesc& hd, int nbuf, int direct) {
    constexpr int NTW = 2 * NBW * 8;
    int max_cp = 0;
    a.kt_tiles = 0;
    for (int i = 0; i < kMaxSrc; ++i) a.ntiles_k[i] = 0;
    for (int i = 0; i < hd.nsrc; ++i) { max_cp = std::max(max_cp, hd.src[i].cp); a.ntiles_k[i] = (hd.src[i].cp + kWgK - 1) / kWgK; a.kt_tiles += a.ntiles_k[i]; }
    a.nt_tiles = hd.cols.nplanes * ((hd.cols.gwp + NTW - 1) / NTW);
    const PwWgradSmem L = pw_wgrad_smem(NTW, a.cpo, max_cp, nbuf, direct, a.dr ? hd.NPall : 0);
    if (L.total > kMaxDynSmem) return false;
    auto k = pw_wgrad_kernel<NBW>;
    static bool attr_done = (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), true);
    (void)attr_done;
    a.nbuf = nbuf; a.direct = direct;
    const int gy = a.kt_tiles * a.nt_tiles;
    const int per_sm = std::max(1, std::min(2, (227 * 1024) / (L.total + 1024)));
    const int tps = (a.Rt + kWgR - 1) / kWgR, ntile = kT * tps;
    int gx = std::max(1, num_sms() * per_sm / gy);
    if (gx > ntile) gx = ntile;
    a.tiles_per_cta = (ntile + gx - 1) / gx;
    gx = (ntile + a.tiles_per_cta - 1) / a.tiles_per_cta;
    CDRA_LAUNCH_PDL(k, dim3(gx, gy), dim3(256), L.total, c.stream, a);
    return true;
}

// weight gradient on tcgen05 with the accumulator resident in TMEM (every layer whose [KP x NPall] tile fits 512 columns)
template <int R, int NT = 512>
inline bool try_pw_wgrad_tc(const RunCtx& c, PwBwdArgs& a, const PwDesc& hd, int min_ring) {
    int sum = 0;
    for (int i = 0; i < hd.nsrc; ++i) sum += hd.src[i].cp;
    const PwWgTcSmem L0 = pw_wgrad_tc_smem(R, hd.KP, hd.NPall, hd.cols.nplanes, a.cpo, sum, 0);
    if (L0.np > 256 || L0.mb * L0.np > 512) return false;
    // one CTA per SM; whatever shared memory the staging tiles leave goes to the TMA ring (bytes in flight hide the HBM latency)
    const int nbuf = std::min(8, (kMaxDynSmem - L0.total) / L0.raw_stride);
    if (nbuf < min_ring) return false;
    const PwWgTcSmem L = pw_wgrad_tc_smem(R, hd.KP, hd.NPall, hd.cols.nplanes, a.cpo, sum, nbuf);
    auto k = pw_wgrad_tc_kernel<R, NT>;
    static bool attr_done = (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), true);
    (void)attr_done;
    a.nbuf = nbuf; a.direct = 0;
    const int tps = (a.Rt + R - 1) / R, ntile = kT * tps;
    int gx = std::min(ntile, num_sms());
    a.tiles_per_cta = (ntile + gx - 1) / gx;
    gx = (ntile + a.tiles_per_cta - 1) / a.tiles_per_cta;
    CDRA_LAUNCH_PDL(k, dim3(gx), dim3(NT), L.total, c.stream, a);
    return true;
}

// the whole backward of one pointwise layer in ONE warp-specialised tcgen05 kernel (v3_pw_bwd.cuh): every layer whose
// weight-gradient accumulator [KP x NPall] plus the double-buffered data-gradient tile fit the 512 TMEM columns
template <int R>
inline bool try_pw_bwd_fused(const RunCtx& c, PwBwdArgs& a, const PwDesc& hd, int min_stages) {
    int cps[kMaxSrc] = {0, 0, 0}, acc[kMaxSrc] = {0, 0, 0};
    for (int i = 0; i < hd.nsrc; ++i) { cps[i] = hd.src[i].cp; acc[i] = hd.src[i].accumulate; }
    const int x1cp = a.x1 ? a.x1cp : 0;
    const PwBfSmem L0 = pw_bf_smem(R, hd.nsrc, cps, acc, hd.NPall, hd.cols.nplanes, a.cpo, x1cp, 0);
    if (L0.mbk != 1 || L0.np16 > 256 || L0.cols_dw + 2 * R > 512 || x1cp > kBfEpilogueThreads) return false;
    const int nstage = std::min(kBfMaxStages, (kMaxDynSmem - 1024 - L0.ring) / L0.stage_bytes);
    if (nstage < min_stages) return false;
    const PwBfSmem L = pw_bf_smem(R, hd.nsrc, cps, acc, hd.NPall, hd.cols.nplanes, a.cpo, x1cp, nstage);
    if (L.total > kMaxDynSmem) return false;
    auto k = pw_bwd_fused_kernel<R>;
    static bool attr_done = (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), true);
    (void)attr_done;
    a.nbuf = nstage; a.direct = 0;
    { static const int only_r = getenv("CDRA_TIMELINE_R") ? atoi(getenv("CDRA_TIMELINE_R")) : 0; if (only_r && only_r != R) a.timeline = 0; }
    const int tps = (a.Rt + R - 1) / R, ntile = kT * tps;
    int gx = std::min(ntile, num_sms());
    a.tiles_per_cta = (ntile + gx - 1) / gx;
    gx = (ntile + a.tiles_per_cta - 1) / a.tiles_per_cta;
    CDRA_LAUNCH_PDL(k, dim3(gx), dim3(kBfThreads), L.total, c.stream, a);
    return true;
}

inline bool try_pwg_dgrad(const RunCtx& c, PwBwdArgs& a, const PwDesc& hd) {
    if (hd.KP < pwg_min_k_bwd() || hd.NPall > 768) return false;
    for (int i = 0; i < hd.nsrc; ++i)            // the vectorised sums kernel covers whole 8-slot chunks
        if (hd.src[i].bsum && hd.src[i].sum_hi > hd.src[i].sum_lo && ((hd.src[i].sum_lo & 7) || ((hd.src[i].sum_hi - hd.src[i].sum_lo) & 7) || hd.src[i].sum_hi - hd.src[i].sum_lo > 256)) return false;
    a.nblk = 0;
    int off = 0;
    for (int i = 0; i < hd.nsrc; ++i) {
        for (int s0 = 0; s0 < hd.src[i].cp; s0 += 128) {
            if (a.nblk == kGMaxBlk) return false;
            a.blk[a.nblk++] = GBlock{off + s0, std::min(128, hd.src[i].cp - s0), i, s0};
        }
        off += hd.src[i].cp;
    }
    int ares, nstage, smem;
    if (!pwg_plan(hd.NPall, ((hd.NPall + 63) & ~63) * 16, 16384, ares, nstage, smem)) return false;
    a.nbuf = nstage;
    int gx; pwg_grid(a.Rt, a.nblk, gx, a.tiles_per_cta);
    const int depth = std::max(1, std::min(4, nstage - 2));
    auto launch = [&](auto k) {
        max_smem_once(k);
        CDRA_LAUNCH_PDL(k, dim3(gx, a.nblk), dim3(kGThreads), smem, c.stream, a, ares);
    };
    if (depth == 4) launch(pwg_dgrad_kernel<4>); else if (depth == 3) launch(pwg_dgrad_kernel<3>);
    else if (depth == 2) launch(pwg_dgrad_kernel<2>); else launch(pwg_dgrad_kernel<1>);
    for (int i = 0; i < hd.nsrc; ++i) {      // BatchNorm-backward sums of the gradients just written
        const PwSrc& Sx = hd.src[i];
        if (!Sx.bsum || Sx.sum_hi <= Sx.sum_lo) continue;
        BsumArgs q; q.x = Sx.data; q.dx = Sx.grad; q.cp = Sx.cp; q.lo = Sx.sum_lo; q.hi = Sx.sum_hi; q.clamp = Sx.clamp; q.aff = Sx.aff; q.bnp = Sx.bnp;
        static const int bsum_rows = getenv("CDRA_BSUM_ROWS") ? atoi(getenv("CDRA_BSUM_ROWS")) : 128;
        q.bsum = Sx.bsum; q.Rt = a.Rt; q.rows_per_cta = bsum_rows;
        prof_bytes(4.0 * a.Rt * (Sx.sum_hi - Sx.sum_lo) * 2 * 2);
        CDRA_LAUNCH_PDL(bsum_kernel, dim3((a.Rt + q.rows_per_cta - 1) / q.rows_per_cta, kT), dim3(256), 0, c.stream, q);
    }
    if (a.x1) {
        PassBwdArgs q; q.x1 = a.x1; q.dx1 = a.dx1; q.x1cp = a.x1cp; q.x1map = a.x1map; q.x1aff = a.x1aff; q.x1bnp = a.x1bnp; q.x1bsum = a.x1bsum;
        q.x1clamp = a.x1clamp; q.dout[0] = a.dout[0]; q.dout[1] = a.dout[1]; q.cpo = a.cpo; q.ncopy = a.ncopy; q.copy_dst0 = a.copy_dst0; q.Rt = a.Rt;
        q.rows_per_cta = 64;
        prof_bytes(4.0 * a.Rt * 2 * a.ncopy * 2 * 3);
        CDRA_LAUNCH_PDL(pass_bwd_kernel, dim3((a.Rt + q.rows_per_cta - 1) / q.rows_per_cta, kT), dim3(256), 0, c.stream, q);
    }
    return true;
}

// split-K tcgen05 weight gradient over the dR hand-off (v4_pwg.cuh)
inline bool try_pwg_wgrad(const RunCtx& c, const PwBwdArgs& b, const PwDesc& hd) {
    if (hd.KP < pwg_min_k_bwd() || hd.KP > 768 || b.dr == nullptr) return false;
    PwgWgArgs a; memset(&a, 0, sizeof a);
    a.d = b.d; a.Rt = b.Rt; a.dr = b.dr; a.tb[0] = b.tb[0]; a.tb[1] = b.tb[1]; a.cpo = b.cpo;
    int off = 0;
    for (int i = 0; i < hd.nsrc; ++i) {
        for (int s0 = 0; s0 < hd.src[i].cp; s0 += 128) {
            if (a.nblk == kGMaxBlk) return false;
            a.blk[a.nblk++] = GBlock{off + s0, std::min(128, hd.src[i].cp - s0), i, s0};
        }
        off += hd.src[i].cp;
    }
    a.nnb = (hd.NPall + 255) / 256;
    const int nbk = (std::min(256, hd.NPall) + 63) / 64;
    const PwgWgSmem L0 = pwg_wg_smem(hd.KP, nbk, 0);
    int nstage = std::min(kGMaxStages, (kMaxDynSmem - L0.total) / L0.stage_bytes);
    if (nstage < 2) return false;
    const int smem = pwg_wg_smem(hd.KP, nbk, nstage).total;
    a.nbuf = nstage;
    const int gy = a.nblk * a.nnb;
    const int tps = (a.Rt + kWgRows - 1) / kWgRows, ntile = kT * tps;
    const int splits = std::max(1, num_sms() / gy);
    a.tiles_per_cta = (ntile + splits - 1) / splits;
    const int gx = (ntile + a.tiles_per_cta - 1) / a.tiles_per_cta;
    const int depth = std::max(1, std::min(3, nstage - 2));          // two ring stages of slack (see try_pwg_fwd)
    auto launch = [&](auto k) {
        max_smem_once(k);
        CDRA_LAUNCH_PDL(k, dim3(gx, gy), dim3(kGThreads), smem, c.stream, a);
    };
    if (depth == 3) launch(pwg_wgrad_kernel<3>); else if (depth == 2) launch(pwg_wgrad_kernel<2>); else launch(pwg_wgrad_kernel<1>);
    return true;
}

// data gradient (+ pass-through, + BN-backward sums of the inputs) and weight gradient of one GEMM launch
inline void launch_pw_bwd(const RunCtx& c, int di, const PwDesc& hd, PwBwdArgs a) {
    a.d = desc_dev(c, di);
    a.out_clamp = 1;
    a.timeline = use_timeline();
    if (use_tc() && use_fused()) {
        double fb = 0;
        for (int i = 0; i < hd.nsrc; ++i) fb += 4.0 * a.Rt * (hd.src[i].map.n0 + hd.src[i].map.n1) * 2 * (2 + hd.src[i].accumulate);    // raw src read, d src written (+ read)
        fb += 4.0 * a.Rt * hd.cols.nplanes * (hd.cols.seg0n + hd.cols.seg1n + a.ncopy) * 2 * 2;                                       // d out, out read
        if (a.x1) fb += 4.0 * a.Rt * 2 * a.ncopy * 2 * 2;
        prof_bytes(fb);
        static const int min64 = getenv("CDRA_BF_MIN64") ? atoi(getenv("CDRA_BF_MIN64")) : 4;      // ring depth below which 32-row tiles are preferred
        if (try_pw_bwd_fused<64>(c, a, hd, min64) || try_pw_bwd_fused<32>(c, a, hd, 3) || try_pw_bwd_fused<64>(c, a, hd, 2) || try_pw_bwd_fused<32>(c, a, hd, 2)) return;
    }
    const int tc_np = (hd.NPall + 15) & ~15, tc_mb = (hd.KP + 127) / 128;
    const bool tc = use_tc() && tc_np <= 256 && tc_mb * tc_np <= 512;       // the [KP x NPall] accumulator fits the SM's TMEM
    a.dr = use_tc() ? (bf16*)(c.ws + c.p->v2.dr_scratch) : nullptr;     // every layer hands dR over; the mma.sync weight gradient takes it too
    int max_cp = 0;
    for (int i = 0; i < hd.nsrc; ++i) max_cp = std::max(max_cp, hd.src[i].cp);
    double bytes = 0;
    for (int i = 0; i < hd.nsrc; ++i) bytes += 4.0 * a.Rt * (hd.src[i].map.n0 + hd.src[i].map.n1) * 2 * 2;          // raw src read, d src written
    bytes += 4.0 * a.Rt * hd.cols.nplanes * (hd.cols.seg0n + hd.cols.seg1n + a.ncopy) * 2 * 2;                      // d out, out read
    if (a.x1) bytes += 4.0 * a.Rt * 2 * a.ncopy * 2 * 2;
    prof_bytes(bytes);
    bool ok, pwg = false;
    if (use_tc() && use_pwg() && a.x1cp <= 256 && try_pwg_dgrad(c, a, hd)) ok = pwg = true;
    else if (max_cp <= 64)
        ok = try_pw_dgrad<64, 4, 2, 1, 4>(c, a, hd, 2, 0, 2) || try_pw_dgrad<64, 4, 2, 1, 4>(c, a, hd, 1, 0, 2) ||
             try_pw_dgrad<64, 4, 2, 1, 4>(c, a, hd, 2, 0, 1) || try_pw_dgrad<64, 4, 2, 1, 4>(c, a, hd, 1, 0, 1) ||
             try_pw_dgrad<32, 2, 4, 1, 2>(c, a, hd, 1, 0, 1) || try_pw_dgrad<32, 2, 4, 1, 2>(c, a, hd, 1, 1, 1);
    else
        ok = try_pw_dgrad<64, 4, 2, 1, 8>(c, a, hd, 2, 0, 2) || try_pw_dgrad<64, 4, 2, 1, 8>(c, a, hd, 1, 0, 2) ||
             try_pw_dgrad<64, 4, 2, 1, 8>(c, a, hd, 2, 0, 1) || try_pw_dgrad<64, 4, 4, 1, 4>(c, a, hd, 2, 0, 1) || try_pw_dgrad<64, 4, 4, 1, 4>(c, a, hd, 1, 0, 1) ||      // 16 warps, one CTA per SM try_pw_dgrad<64, 4, 2, 1, 8>(c, a, hd, 1, 0, 1) ||
             try_pw_dgrad<32, 2, 4, 1, 4>(c, a, hd, 1, 0, 1) || try_pw_dgrad<32, 2, 4, 1, 4>(c, a, hd, 1, 1, 1) ||
             try_pw_dgrad<32, 2, 4, 1, 2>(c, a, hd, 1, 1, 1);
    if (!ok) fprintf(stderr, "libcdra: no pw_dgrad configuration fits (KP=%d NPall=%d)\n", hd.KP, hd.NPall);
    bytes = 0;
    for (int i = 0; i < hd.nsrc; ++i) bytes += 4.0 * a.Rt * (hd.src[i].map.n0 + hd.src[i].map.n1) * 2;
    bytes += 4.0 * a.Rt * hd.cols.nplanes * (hd.cols.seg0n + hd.cols.seg1n) * 2 * 2;
    prof_bytes(bytes);
    if (pwg && use_pwg_wgrad() && try_pwg_wgrad(c, a, hd)) return;
    if (tc && (try_pw_wgrad_tc<128>(c, a, hd, 3) || try_pw_wgrad_tc<64>(c, a, hd, 4) || try_pw_wgrad_tc<32>(c, a, hd, 4) || try_pw_wgrad_tc<16>(c, a, hd, 3) ||
                     try_pw_wgrad_tc<32>(c, a, hd, 2) || try_pw_wgrad_tc<16>(c, a, hd, 2))) return;
    if (hd.cols.gwp <= 64)
        ok = try_pw_wgrad<4>(c, a, hd, 2, 0) || try_pw_wgrad<4>(c, a, hd, 1, 0) || try_pw_wgrad<4>(c, a, hd, 1, 1);
    else
        ok = try_pw_wgrad<8>(c, a, hd, 2, 0) || try_pw_wgrad<8>(c, a, hd, 1, 0) || try_pw_wgrad<8>(c, a, hd, 1, 1);
    if (!ok) fprintf(stderr, "libcdra: no pw_wgrad configuration fits (KP=%d NPall=%d)\n", hd.KP, hd.NPall);
}

inline void launch_dw_bwd(const RunCtx& c, const BnConv& l, const V2Tensor& in, bool clamp, int kbase, const V2Tensor& out,
                          const Unit& u, bool accumulate, bool final_grad) {
    DwArgs a; memset(&a, 0, sizeof a);
    a.in = (const bf16*)(c.ws + in.data); a.aff = in.has_bn ? (const float2*)(c.ws + in.aff) : nullptr;
    a.bnp = in.has_bn ? (const float2*)(c.ws + in.bnp) : nullptr; a.clamp = clamp ? 1 : 0;
    a.din = (bf16*)(c.ws + in.grad); a.accumulate = accumulate ? 1 : 0;
    a.in_bsum = (final_grad && in.has_bn) ? (double2*)(c.ws + in.bsum) : nullptr; a.in_sum_lo = 0; a.in_sum_hi = in.cp;
    a.cp = in.cp; a.map = SlotMap{in.n0, in.n0p, in.n1}; a.kbase = kbase;
    a.B = c.p->B; a.Hi = u.Hi; a.Wi = u.Wi; a.Ho = u.Ho; a.Wo = u.Wo; a.stride = u.stride; a.pad_t = u.pad_t; a.pad_l = u.pad_l;
    a.L = layer_of(c, l);
    a.out = (bf16*)(c.ws + out.data); a.dout = (const bf16*)(c.ws + out.grad); a.tb = tables_of(c, out);
    a.training = 1;
    typedef void (*KernelT)(const DwArgs);
    KernelT k = nullptr;
    switch (in.cp * 10 + u.stride) {
        case 641: k = dw_bwd_kernel<64, 1>; break;    case 1201: k = dw_bwd_kernel<120, 1>; break;  case 2321: k = dw_bwd_kernel<232, 1>; break;
        case 242: k = dw_bwd_kernel<24, 2>; break;    case 642: k = dw_bwd_kernel<64, 2>; break;
        case 1202: k = dw_bwd_kernel<120, 2>; break;  case 2322: k = dw_bwd_kernel<232, 2>; break;
    }
    if (!k) { fprintf(stderr, "libcdra: no depthwise kernel for cp=%d stride=%d\n", in.cp, u.stride); return; }
    a.nbuf = 2;
    if (!dw_pick_band(in.cp, u, true, kMaxDynSmem, a.band_rows, a.nbands)) { fprintf(stderr, "libcdra: dw_bwd band does not fit in shared memory (cp=%d)\n", in.cp); return; }
    const DwSmem L = dw_smem(in.cp, u.Hi, u.Wi, u.Ho, u.Wo, a.nbuf, true, a.band_rows, u.stride);
    max_smem_once(k);
    const int nitems = kT * a.B * a.nbands;
    int gx = std::min(nitems, num_sms());
    a.frames_per_cta = (nitems + gx - 1) / gx;
    gx = (nitems + a.frames_per_cta - 1) / a.frames_per_cta;
    prof_bytes(4.0 * a.B * (2.0 * u.Hi * u.Wi + 2.0 * u.Ho * u.Wo) * (in.n0 + in.n1) * 2);
    CDRA_LAUNCH_PDL(k, dim3(gx), dim3(kDwThreads), L.total, c.stream, a);
}

inline void tower_backward(const RunCtx& c) {
    const Plan& p = *c.p; const V2Plan& v = p.v2;
    const int nu = (int)v.u.size();
    upload_descs(c);
    const PwDesc* host = (const PwDesc*)v.host_descs;
    {   // global average pool + head conv
        const V2Tensor& th = v.t[v.head];
        GapArgs g; memset(&g, 0, sizeof g);
        g.in = (const bf16*)(c.ws + th.data); g.aff = (const float2*)(c.ws + th.aff); g.bnp = (const float2*)(c.ws + th.bnp);
        g.cp = th.cp; g.C = th.n0; g.HW = th.H * th.W; g.F = kT * p.B; g.B = p.B;
        g.dgap = (const float*)(c.ws + p.dgap); g.dout = (bf16*)(c.ws + th.grad); g.bsum = (double2*)(c.ws + th.bsum);
        const int fpb = 4;
        prof_bytes((double)g.F * g.HW * g.C * 2 * 2);
        CDRA_LAUNCH(gap_bwd_kernel, dim3(cdiv(p.B, fpb), kT), dim3(256), th.cp * 8, c.stream, g, fpb);
        PwBwdArgs a; memset(&a, 0, sizeof a);
        a.Rt = th.Rt; a.out[0] = a.out[1] = (const bf16*)(c.ws + th.data); a.dout[0] = a.dout[1] = (const bf16*)(c.ws + th.grad);
        a.cpo = th.cp; a.tb[0] = a.tb[1] = tables_of(c, th);
        launch_pw_bwd(c, 2 * nu, host[2 * nu], a);
    }
    for (int ui = nu - 1; ui >= 0; --ui) {
        const V2Unit& u = v.u[ui]; const Unit& un = p.units[ui];
        const V2Tensor& r1 = v.t[u.r1]; const V2Tensor& r2 = v.t[u.r2];
        const V2Tensor& oA = v.t[u.outA]; const V2Tensor& oB = v.t[u.outB];
        {   // tail: pw2 (+ shortcut pw | pass-through)
            PwBwdArgs a; memset(&a, 0, sizeof a);
            a.Rt = oA.Rt; a.out[0] = (const bf16*)(c.ws + oA.data); a.out[1] = (const bf16*)(c.ws + oB.data);
            a.dout[0] = (const bf16*)(c.ws + oA.grad); a.dout[1] = (const bf16*)(c.ws + oB.grad);
            a.cpo = oA.cp; a.tb[0] = tables_of(c, oA); a.tb[1] = tables_of(c, oB);
            if (un.stride == 1) {
                const V2Tensor& x1 = v.t[u.inA];
                a.x1 = (const bf16*)(c.ws + x1.data); a.dx1 = (bf16*)(c.ws + x1.grad); a.x1cp = x1.cp; a.x1map = SlotMap{x1.n0, x1.n0p, x1.n1};
                a.x1aff = (const float2*)(c.ws + x1.aff); a.x1bnp = (const float2*)(c.ws + x1.bnp); a.x1bsum = (double2*)(c.ws + x1.bsum);
                a.x1clamp = 1; a.ncopy = oA.n1; a.copy_dst0 = oA.n0p;
            }
            launch_pw_bwd(c, 2 * ui + 1, host[2 * ui + 1], a);
        }
        launch_dw_bwd(c, un.dw, r1, true, 0, r2, un, false, true);
        if (un.stride == 2) {   // shortcut depthwise first (plain write), pw1 then adds its share and finalises the sums
            const bool in_clamp = u.inB >= 0;
            launch_dw_bwd(c, un.scdw, v.t[u.inA], in_clamp, 0, v.t[u.rsA], un, false, false);
            if (u.inB >= 0) launch_dw_bwd(c, un.scdw, v.t[u.inB], in_clamp, v.t[u.inA].C(), v.t[u.rsB], un, false, false);
        }
        {   // pw1
            PwBwdArgs a; memset(&a, 0, sizeof a);
            a.Rt = r1.Rt; a.out[0] = a.out[1] = (const bf16*)(c.ws + r1.data); a.dout[0] = a.dout[1] = (const bf16*)(c.ws + r1.grad);
            a.cpo = r1.cp; a.tb[0] = a.tb[1] = tables_of(c, r1);
            launch_pw_bwd(c, 2 * ui, host[2 * ui], a);
        }
    }
}


// ================================================================================================ stem (uint8 frames)
inline StemGeom stem_geom(const Plan& p) {
    StemGeom g; g.B = p.B; g.H = p.H; g.W = p.W; g.W3 = 3 * p.W; g.Hs = p.Hs; g.Ws = p.Ws; g.Hp = p.Hp; g.Wp = p.Wp;
    g.pad_t = p.pool_pad_t; g.pad_l = p.pool_pad_l;
    return g;
}
// rows per band: the largest count whose shared-memory footprint stays under `limit`, evened out over the bands
template <typename F>
inline int band_rows(int rows, int limit, F smem_of) {
    int hb = 1;
    while (hb < rows && smem_of(hb + 1) <= limit) ++hb;
    const int nb = (rows + hb - 1) / hb;
    return (rows + nb - 1) / nb;
}
inline void persistent_grid(int nunits, int ctas_per_sm, int& gx, int& per_cta) {
    gx = std::min(nunits, num_sms() * ctas_per_sm);
    per_cta = (nunits + gx - 1) / gx;
    gx = (nunits + per_cta - 1) / per_cta;
}

inline void stem_forward(const RunCtx& c, const uint8_t* image) {
    const Plan& p = *c.p; const StemGeom g = stem_geom(p);
    const WsTensor& ts = p.tensors[p.t_stem]; const WsTensor& tp = p.tensors[p.t_pool];
    const int limit = 108 * 1024;
    {   // conv 3x3 s2 (+ sums of the stored values)                         core/architectures.py:159
        StemFwdArgs a; memset(&a, 0, sizeof a);
        a.img = image; a.g = g; a.w = c.params + p.stem.w; a.bias = c.params + p.stem.b;
        a.out = (bf16*)(c.ws + ts.data); a.fst = (double2*)(c.ws + ts.fst); a.training = c.training;
        a.HB = band_rows(g.Hs, limit, [&](int hb) { return 2 * hb * g.W3 + 6 * g.Ws >= 65536 ? (1 << 30) : stem_fwd_smem(hb, g.Ws, g.W3).total; });
        a.nbands = (g.Hs + a.HB - 1) / a.HB;
        int gx; persistent_grid(kT * g.B * a.nbands, 2, gx, a.units_per_cta);
        const int smem = stem_fwd_smem(a.HB, g.Ws, g.W3).total;
        static bool attr = (cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), true); (void)attr;
        prof_bytes(4.0 * g.B * ((double)g.H * g.W3 + (double)g.Hs * g.Ws * kSC * 2));
        CDRA_LAUNCH(stem_fwd_kernel, dim3(gx), dim3(kStemThreads), smem, c.stream, a);
        launch_bn_finalize(c, p.stem, ts, ColMap{kSC, 0, 0, 0}, (double)ts.Rt);
    }
    {   // BN + ReLU6, max pool 3x3 s2 SAME (+ winner positions)              :160-161
        PoolFwdArgs a; memset(&a, 0, sizeof a);
        a.stem = (const bf16*)(c.ws + ts.data); a.aff = (const float2*)(c.ws + ts.aff); a.g = g;
        a.pool = (bf16*)(c.ws + tp.data); a.idx = (uint8_t*)(c.ws + p.v2.stem_idx);
        a.PB = band_rows(g.Hp, 44 * 1024, [&](int pb) { return pool_fwd_smem(pb, g.Ws).total; });
        a.nbands = (g.Hp + a.PB - 1) / a.PB;
        int gx; persistent_grid(kT * g.B * a.nbands, 4, gx, a.units_per_cta);
        const int smem = pool_fwd_smem(a.PB, g.Ws).total;
        static bool attr = (cudaFuncSetAttribute(pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), true); (void)attr;
        prof_bytes(4.0 * g.B * ((double)g.Hs * g.Ws + (double)g.Hp * g.Wp) * kSC * 2);
        CDRA_LAUNCH(pool_fwd_kernel, dim3(gx), dim3(kStemThreads), smem, c.stream, a);
    }
}

inline void stem_backward(const RunCtx& c, const uint8_t* image) {
    const Plan& p = *c.p; const StemGeom g = stem_geom(p);
    const WsTensor& ts = p.tensors[p.t_stem]; const WsTensor& tp = p.tensors[p.t_pool];
    StemBwdArgs a; memset(&a, 0, sizeof a);
    a.img = image; a.g = g; a.stem = (const bf16*)(c.ws + ts.data); a.dpool = (const bf16*)(c.ws + tp.grad);
    a.idx = (const uint8_t*)(c.ws + p.v2.stem_idx); a.aff = (const float2*)(c.ws + ts.aff); a.bnp = (const float2*)(c.ws + ts.bnp);
    a.gacc = (double*)(c.ws + p.v2.stem_gacc);
    a.HB = band_rows(g.Hs, 108 * 1024, [&](int hb) { return 2 * hb * g.W3 + 6 * g.Ws >= 65536 ? (1 << 30) : stem_bwd_smem(hb, g.Ws, g.Wp, g.W3).total; });
    a.nbands = (g.Hs + a.HB - 1) / a.HB;
    int gx; persistent_grid(kT * g.B * a.nbands, 2, gx, a.units_per_cta);
    const int smem = stem_bwd_smem(a.HB, g.Ws, g.Wp, g.W3).total;
    static bool attr = (cudaFuncSetAttribute(stem_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), true); (void)attr;
    prof_bytes(4.0 * g.B * ((double)g.H * g.W3 + (double)g.Hs * g.Ws * kSC * 2 + (double)g.Hp * g.Wp * kSC * 2));
    CDRA_LAUNCH(stem_bwd_kernel, dim3(gx), dim3(kStemThreads), smem, c.stream, a);
    StemFinArgs f; f.gacc = a.gacc; f.aff = a.aff; f.dw = c.grads + p.stem.w; f.dgamma = c.grads + p.stem.g; f.dbeta = c.grads + p.stem.be;
    f.n = (double)ts.Rt;
    CDRA_LAUNCH(stem_bwd_finish_kernel, dim3(1), dim3(kStemThreads), 0, c.stream, f);
}

// ---- logical-layout export of a tower tensor (or its gradient) for the parity taps: out [4B][H][W][C] fp32.
// Unit outputs ("...out") are assembled from both planes in the reference's channel order.
struct ExportArgs { const bf16* a; const bf16* b; int cp; SlotMap map; long long rows; int C; float* out; };
__global__ void __launch_bounds__(256) export_kernel(const ExportArgs e) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= e.rows * e.C) return;
    const long long r = i / e.C; int ch = (int)(i - r * e.C);
    const bf16* src = e.a;
    const int half = e.map.n0 + e.map.n1;
    if (e.b && ch >= half) { src = e.b; ch -= half; }
    e.out[i] = __bfloat162float(src[r * e.cp + logical_slot(e.map, ch)]);
}
// BatchNorm tables of a tower tensor in the logical channel order: out [4][C][2] fp32.  which 0: aff (scale, shift),
// 1: bnp (mean, inv_std), 2: bsum (sum dz, sum dz * xhat; fp64 in the workspace)
struct ExportTabArgs { const void* a; const void* b; int cp; SlotMap map; int C, half, which; float* out; };
__global__ void __launch_bounds__(256) export_tab_kernel(const ExportTabArgs e) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= kT * e.C) return;
    const int t = i / e.C; int ch = i - t * e.C;
    const void* src = e.a;
    if (e.b && ch >= e.half) { src = e.b; ch -= e.half; }
    const size_t idx = (size_t)t * e.cp + logical_slot(e.map, ch);
    float2 v;
    if (e.which == 2) { const double2 d = reinterpret_cast<const double2*>(src)[idx]; v = make_float2((float)d.x, (float)d.y); }
    else v = reinterpret_cast<const float2*>(src)[idx];
    e.out[2 * i] = v.x; e.out[2 * i + 1] = v.y;
}

inline bool export_tensor(const Plan& p, const char* name_c, char* ws, float* out, int32_t dims[4], cudaStream_t st) {
    const V2Plan& v = p.v2;
    std::string n(name_c);
    bool grad = false;
    int tab = -1;
    if (n.rfind("grad:", 0) == 0) { grad = true; n = n.substr(5); }
    else if (n.rfind("aff:", 0) == 0) { tab = 0; n = n.substr(4); }
    else if (n.rfind("bnp:", 0) == 0) { tab = 1; n = n.substr(4); }
    else if (n.rfind("bsum:", 0) == 0) { tab = 2; n = n.substr(5); }
    if (tab >= 0 && tab < 2 && (n == "tower.stem")) {          // the stem keeps the legacy (unpadded) table layout
        const WsTensor& t = p.tensors[p.t_stem];
        ExportTabArgs e; e.a = ws + (tab == 0 ? t.aff : t.bnp); e.b = nullptr; e.cp = t.C; e.map = SlotMap{t.C, t.C, 0};
        e.C = t.C; e.half = t.C; e.which = tab; e.out = out;
        if (dims) { dims[0] = kT; dims[1] = t.C; dims[2] = 2; dims[3] = 1; }
        if (out) { CDRA_LAUNCH(export_tab_kernel, dim3(cdiv(kT * e.C, 256)), dim3(256), 0, st, e); }
        return true;
    }
    int ia = -1, ib = -1;
    auto it = v.index.find(n);
    if (it != v.index.end()) ia = it->second;
    else if (n.size() > 4 && n.substr(n.size() - 4) == ".out") {
        auto a = v.index.find(n + "A"), b = v.index.find(n + "B");
        if (a != v.index.end() && b != v.index.end()) { ia = a->second; ib = b->second; }
    } else if (n.size() > 5 && n.substr(n.size() - 5) == ".scdw") {
        auto a = v.index.find(n + "A"), b = v.index.find(n + "B");
        if (a != v.index.end()) { ia = a->second; if (b != v.index.end()) ib = b->second; }
    }
    if (ia < 0) return false;
    const V2Tensor& ta = v.t[ia];
    if (tab >= 0) {
        if (!ta.has_bn) return false;
        auto sel = [&](const V2Tensor& t) { return (const void*)(ws + (tab == 0 ? t.aff : tab == 1 ? t.bnp : t.bsum)); };
        ExportTabArgs e; e.a = sel(ta); e.b = ib >= 0 ? sel(v.t[ib]) : nullptr; e.cp = ta.cp; e.map = SlotMap{ta.n0, ta.n0p, ta.n1};
        e.half = ta.C(); e.C = ta.C() * (ib >= 0 ? 2 : 1); e.which = tab; e.out = out;
        if (dims) { dims[0] = kT; dims[1] = e.C; dims[2] = 2; dims[3] = 1; }
        if (out) { CDRA_LAUNCH(export_tab_kernel, dim3(cdiv(kT * e.C, 256)), dim3(256), 0, st, e); }
        return true;
    }
    ExportArgs e;
    e.a = (const bf16*)(ws + (grad ? ta.grad : ta.data)); e.b = ib >= 0 ? (const bf16*)(ws + (grad ? v.t[ib].grad : v.t[ib].data)) : nullptr;
    e.cp = ta.cp; e.map = SlotMap{ta.n0, ta.n0p, ta.n1}; e.rows = (long long)4 * ta.Rt; e.C = ta.C() * (ib >= 0 ? 2 : 1); e.out = out;
    if (dims) { dims[0] = 4 * p.B; dims[1] = ta.H; dims[2] = ta.W; dims[3] = e.C; }
    if (out) { CDRA_LAUNCH(export_kernel, dim3(cdiv(e.rows * e.C, 256)), dim3(256), 0, st, e); }
    return true;
}

}  // namespace v2
}  // namespace cdra
#endif
