// Second-generation versions of tower kernels that replaced slow first implementations.
#pragma once
#include "tower_bwd.cuh"

namespace cdra {

// --------------------------------------------------------------------------- maxpool backward (first max wins)
// The forward pool output P is max(window); a stem pixel receives the window's gradient iff its activated
// value equals P and no earlier pixel of the window (row-major scan, TF/Eigen MaxPoolGrad order) does.
// One 2-channel load of P per window instead of re-scanning 9 taps; earlier taps are only inspected for
// genuine candidates.
template <typename T>
struct PoolBwd2Args {
    ActView in;                 // stem raw (+affine, ReLU6)
    const T* pool;              // forward pool output [kT*B*Ho*Wo][C]
    const T* dpool;
    T* dstem;
    int B, Hi, Wi, Ho, Wo, C, pad_t, pad_l;
};

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) pool_bwd2_kernel(PoolBwd2Args<T> a) {
    const int t = blockIdx.y;
    const int hiw = a.Hi * a.Wi, CP = a.C >> 1;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)a.B * hiw * CP) return;
    const int c = (int)(idx % CP) * 2;
    const long long p = idx / CP;
    const int b = (int)(p / hiw), r = (int)(p - (long long)b * hiw), y = r / a.Wi, x = r - y * a.Wi;
    const T* base = (const T*)a.in.data + ((size_t)(t * a.B + b) * hiw) * a.in.ld + a.in.coff + c;
    float2 f0 = make_float2(1.f, 0.f), f1 = make_float2(1.f, 0.f);
    if (a.in.aff) { f0 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c]; f1 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c + 1]; }
    auto act = [&](int yy, int xx) {
        float2 v = ld2(base + ((size_t)yy * a.Wi + xx) * a.in.ld);
        v.x = fmaf(v.x, f0.x, f0.y); v.y = fmaf(v.y, f1.x, f1.y);
        if (a.in.clamp) { v.x = relu6f(v.x); v.y = relu6f(v.y); }
        return v;
    };
    const float2 mine = act(y, x);
    // the pool output was stored in T: compare in the stored precision
    const float m0 = rnd(mine.x, (const T*)nullptr), m1 = rnd(mine.y, (const T*)nullptr);
    float g0 = 0.f, g1 = 0.f;
    const int oy_lo = max(0, (y + a.pad_t - 1) / 2), oy_hi = min(a.Ho - 1, (y + a.pad_t) / 2);
    const int ox_lo = max(0, (x + a.pad_l - 1) / 2), ox_hi = min(a.Wo - 1, (x + a.pad_l) / 2);
    for (int oy = oy_lo; oy <= oy_hi; ++oy)
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
            const size_t o = (((size_t)(t * a.B + b) * a.Ho + oy) * a.Wo + ox) * a.C + c;
            const float2 pv = ld2(a.pool + o);
            bool w0 = (m0 == pv.x), w1 = (m1 == pv.y);
            if (!w0 && !w1) continue;
            // earlier taps of this window (row-major) with the same value take precedence
            const int y0 = oy * 2 - a.pad_t, x0 = ox * 2 - a.pad_l;
            for (int iy = max(y0, 0); iy <= y && (w0 || w1); ++iy) {
                const int xe = (iy < y) ? min(x0 + 2, a.Wi - 1) : x - 1;
                for (int ix = max(x0, 0); ix <= xe; ++ix) {
                    const float2 v = act(iy, ix);
                    if (rnd(v.x, (const T*)nullptr) == pv.x) w0 = false;
                    if (rnd(v.y, (const T*)nullptr) == pv.y) w1 = false;
                }
            }
            if (w0 || w1) {
                const float2 d = ld2(a.dpool + o);
                if (w0) g0 += d.x;
                if (w1) g1 += d.y;
            }
        }
    st2(a.dstem + ((size_t)t * a.B * hiw + p) * a.C + c, make_float2(g0, g1));
}

// --------------------------------------------------------------------------- inference-mode BatchNorm affine
// training=False (CARLANetwork.dynamics_predict, core/networks.py:206-208): every conv's per-(slice, channel)
// (scale, shift) comes from the moving statistics; one launch fills the tables of up to kEvalMax layers.
struct EvalAffLayer { const float* gamma; const float* beta; const float* mm; const float* mv; float2* aff; float2* bnp; int ld; ColMap cm; };
constexpr int kEvalMax = 28;
struct EvalAffArgs { EvalAffLayer l[kEvalMax]; int n; };

CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) eval_affine_kernel(EvalAffArgs a) {
    const EvalAffLayer& L = a.l[blockIdx.x];
    for (int j = threadIdx.x; j < L.cm.n; j += 256) {
        const int c = colmap_c(L.cm, j), w = colmap_w(L.cm, j);
        const float inv = (float)(1.0 / sqrt((double)L.mv[w] + (double)kBnEps));
        const float scale = L.gamma[w] * inv;
        for (int t = 0; t < kT; ++t) {
            L.aff[(size_t)t * L.ld + c] = make_float2(scale, L.beta[w] - L.mm[w] * scale);
            L.bnp[(size_t)t * L.ld + c] = make_float2(L.mm[w], inv);
        }
    }
}

}  // namespace cdra
