// Second-generation versions of tower kernels that replaced slow first implementations.
#pragma once
#include "tower_fwd.cuh"
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
    // grid = (Hi, B, kT): the row is the block index, so no per-thread index divisions by runtime values
    const int t = blockIdx.z, b = blockIdx.y, y = blockIdx.x;
    const int hiw = a.Hi * a.Wi;
    constexpr int CP = kStemC / 2;
  for (int i = threadIdx.x; i < a.Wi * CP; i += 256) {
    const int x = i / CP, c = (i - x * CP) * 2;
    const long long p = (long long)b * hiw + (long long)y * a.Wi + x;
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
}

// --------------------------------------------------------------------------- depthwise 3x3, row-sweep versions
// One thread owns a channel pair and sweeps one output row with a 3x3 register window: S new columns are
// loaded per output pixel (3 loads for stride 1 instead of 9), no per-pixel index divisions.
template <int S, typename LoadF, typename EmitF>
CDRA_DEV void dw_row_sweep(int Wo, int pad_l, LoadF ld, EmitF emit) {
    float2 win[3][3], nxt[3][S];          // nxt: the S new columns of the next step, loaded one step ahead
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = S; kx < 3; ++kx) win[ky][kx] = ld(ky, -S - pad_l + kx);
#pragma unroll
        for (int j = 0; j < S; ++j) nxt[ky][j] = ld(ky, -pad_l + 3 - S + j);
    }
    for (int ox = 0; ox < Wo; ++ox) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3 - S; ++kx) win[ky][kx] = win[ky][kx + S];
#pragma unroll
            for (int j = 0; j < S; ++j) win[ky][3 - S + j] = nxt[ky][j];
        }
        if (ox + 1 < Wo) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int j = 0; j < S; ++j) nxt[ky][j] = ld(ky, (ox + 1) * S - pad_l + 3 - S + j);
        }
        emit(ox, win);
    }
}

struct DwRow { int c, b, y; bool active; };
CDRA_DEV DwRow dw_row_of(int C, int rows_per_image, int nrows, int tid, int block) {
    const DwLanes L = dw_lanes(C, tid);
    DwRow r;
    const int row = block * L.lanes_r + L.rl;
    r.active = L.cl < (C >> 1) && L.rl < L.lanes_r && row < nrows;
    r.c = L.cl * 2; r.b = row / rows_per_image; r.y = row - r.b * rows_per_image;
    return r;
}

template <typename T, int S>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) dw_fwd_row_kernel(DwArgs<T> a) {
    CDRA_SHARED float s_sum[kDwMaxC], s_sq[kDwMaxC];
    const int tid = threadIdx.x, t = blockIdx.y;
    for (int i = tid; i < a.C; i += 256) { s_sum[i] = 0.f; s_sq[i] = 0.f; }
    __syncthreads();
    const DwRow R = dw_row_of(a.C, a.Ho, a.B * a.Ho, tid, blockIdx.x);
    if (R.active) {
        const int c = R.c;
        float w0[9], w1[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { w0[k] = a.w[k * a.C + c]; w1[k] = a.w[k * a.C + c + 1]; }
        const float bias0 = a.bias[c], bias1 = a.bias[c + 1];
        float2 f0 = make_float2(1.f, 0.f), f1 = make_float2(1.f, 0.f);
        if (a.in.aff) { f0 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c]; f1 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c + 1]; }
        const T* img = (const T*)a.in.data + ((size_t)(t * a.B + R.b) * a.Hi * a.Wi) * a.in.ld + a.in.coff + c;
        const int iy0 = R.y * S - a.pad_t;
        const int clampf = a.in.clamp, Wi = a.Wi, Hi = a.Hi, ld = a.in.ld;
        auto load = [&](int ky, int ix) {
            const int iy = iy0 + ky;
            if (iy < 0 || iy >= Hi || ix < 0 || ix >= Wi) return make_float2(0.f, 0.f);
            float2 v = ld2(img + ((size_t)iy * Wi + ix) * ld);
            v.x = fmaf(v.x, f0.x, f0.y); v.y = fmaf(v.y, f1.x, f1.y);
            if (clampf) { v.x = relu6f(v.x); v.y = relu6f(v.y); }
            return v;
        };
        T* orow = a.out + (((size_t)(t * a.B + R.b) * a.Ho + R.y) * a.Wo) * a.C + c;
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
        const int C = a.C;
        dw_row_sweep<S>(a.Wo, a.pad_l, load, [&](int ox, float2 (&win)[3][3]) {
            float a0 = bias0, a1 = bias1;
#pragma unroll
            for (int k = 0; k < 9; ++k) { a0 = fmaf(win[k / 3][k % 3].x, w0[k], a0); a1 = fmaf(win[k / 3][k % 3].y, w1[k], a1); }
            T* dst = orow + (size_t)ox * C;
            st2(dst, make_float2(a0, a1));
            const float v0 = rnd(a0, dst), v1 = rnd(a1, dst);
            s0 += v0; q0 = fmaf(v0, v0, q0); s1 += v1; q1 = fmaf(v1, v1, q1);
        });
        atomicAdd(&s_sum[c], s0); atomicAdd(&s_sq[c], q0);
        atomicAdd(&s_sum[c + 1], s1); atomicAdd(&s_sq[c + 1], q1);
    }
    if (!a.bn.training) return;        // inference (block-uniform)
    __syncthreads();
    for (int i = tid; i < a.C; i += 256) {
        double2* dst = stat_slot(a.tb.fst, a.C, stat_copy(), t, i);
        atomicAdd(&dst->x, (double)s_sum[i]);
        atomicAdd(&dst->y, (double)s_sq[i]);
    }
    const unsigned total_blocks = gridDim.x * gridDim.y;
    if (a.bn.counter != nullptr && last_block_ticket(a.bn.counter, total_blocks)) {
        ColMap cm{a.C, 0, 0, 0};
        bn_finalize(cm, a.tb, a.C, a.bn.gamma, a.bn.beta, a.bn.mov_mean, a.bn.mov_var, (double)a.B * a.Ho * a.Wo,
                    a.bn.unbiased, a.bn.training, 256, tid);
    }
}

// data gradient, stride 1: dX[y][x] = sum_{u,v} dR[y-1+u][x-1+v] * w[2-u][2-v]  (the same sweep with flipped taps)
template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) dw_dgrad_row_kernel(DwBwdArgs<T> a) {
    const int tid = threadIdx.x, t = blockIdx.y;
    const DwRow R = dw_row_of(a.C, a.Hi, a.B * a.Hi, tid, blockIdx.x);
    if (!R.active) return;
    const int c = R.c;
    const double inv_n = 1.0 / ((double)a.B * a.Ho * a.Wo);
    const BnCol b0 = load_bncol(a.tb, a.C, t, c, inv_n), b1 = load_bncol(a.tb, a.C, t, c + 1, inv_n);
    float w0[9], w1[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { w0[k] = a.w[(8 - k) * a.C + c]; w1[k] = a.w[(8 - k) * a.C + c + 1]; }   // flipped
    const size_t img = ((size_t)(t * a.B + R.b) * a.Ho * a.Wo) * a.C + c;
    const int oy0 = R.y - 1, Wo = a.Wo, Ho = a.Ho, C = a.C;
    auto load = [&](int u, int ox) {
        const int oy = oy0 + u;
        if (oy < 0 || oy >= Ho || ox < 0 || ox >= Wo) return make_float2(0.f, 0.f);
        const size_t o = img + ((size_t)oy * Wo + ox) * C;
        const float2 dv = ld2(a.dout + o), rv = ld2(a.out + o);
        return make_float2(make_dr(dv.x, rv.x, b0, 0), make_dr(dv.y, rv.y, b1, 0));
    };
    T* drow = a.dx + (((size_t)(t * a.B + R.b) * a.Hi + R.y) * a.Wi) * a.ldx + a.coffx + c;
    const int ldx = a.ldx, accumulate = a.accumulate;
    dw_row_sweep<1>(a.Wi, 1, load, [&](int ix, float2 (&win)[3][3]) {
        float g0 = 0.f, g1 = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) { g0 = fmaf(win[k / 3][k % 3].x, w0[k], g0); g1 = fmaf(win[k / 3][k % 3].y, w1[k], g1); }
        T* d = drow + (size_t)ix * ldx;
        if (accumulate) { const float2 old = ld2(d); g0 += old.x; g1 += old.y; }
        st2(d, make_float2(g0, g1));
    });
}

template <typename T, int S>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) dw_wgrad_row_kernel(DwBwdArgs<T> a) {
    CDRA_SHARED float s_dw[10][kDwMaxC];        // 9 taps + bias
    const int tid = threadIdx.x, t = blockIdx.y;
    for (int i = tid; i < 10 * kDwMaxC; i += 256) (&s_dw[0][0])[i] = 0.f;
    __syncthreads();
    // persistent over row blocks (grid.x may be smaller than the number of row blocks): the 20 per-thread sums
    // are flushed once per CTA, which keeps the same-address atomic traffic low
    const DwLanes L = dw_lanes(a.C, tid);
    if (L.cl < (a.C >> 1) && L.rl < L.lanes_r) {
        const int c = L.cl * 2, nrows = a.B * a.Ho;
        const double inv_n = 1.0 / ((double)a.B * a.Ho * a.Wo);
        const BnCol b0 = load_bncol(a.tb, a.C, t, c, inv_n), b1 = load_bncol(a.tb, a.C, t, c + 1, inv_n);
        float2 f0 = make_float2(1.f, 0.f), f1 = make_float2(1.f, 0.f);
        if (a.in.aff) { f0 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c]; f1 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c + 1]; }
        float g0[10], g1[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) { g0[k] = 0.f; g1[k] = 0.f; }
      for (int row = blockIdx.x * L.lanes_r + L.rl; row < nrows; row += gridDim.x * L.lanes_r) {
        DwRow R; R.b = row / a.Ho; R.y = row - R.b * a.Ho;
        const T* img = (const T*)a.in.data + ((size_t)(t * a.B + R.b) * a.Hi * a.Wi) * a.in.ld + a.in.coff + c;
        const int iy0 = R.y * S - a.pad_t;
        const int clampf = a.in.clamp, Wi = a.Wi, Hi = a.Hi, ld = a.in.ld, C = a.C;
        auto load = [&](int ky, int ix) {
            const int iy = iy0 + ky;
            if (iy < 0 || iy >= Hi || ix < 0 || ix >= Wi) return make_float2(0.f, 0.f);
            float2 v = ld2(img + ((size_t)iy * Wi + ix) * ld);
            v.x = fmaf(v.x, f0.x, f0.y); v.y = fmaf(v.y, f1.x, f1.y);
            if (clampf) { v.x = relu6f(v.x); v.y = relu6f(v.y); }
            return v;
        };
        const size_t orow = (((size_t)(t * a.B + R.b) * a.Ho + R.y) * a.Wo) * a.C + c;
        dw_row_sweep<S>(a.Wo, a.pad_l, load, [&](int ox, float2 (&win)[3][3]) {
            const size_t o = orow + (size_t)ox * C;
            const float2 dv = ld2(a.dout + o), rv = ld2(a.out + o);
            const float d0 = make_dr(dv.x, rv.x, b0, 0), d1 = make_dr(dv.y, rv.y, b1, 0);
            g0[9] += d0; g1[9] += d1;
#pragma unroll
            for (int k = 0; k < 9; ++k) { g0[k] = fmaf(win[k / 3][k % 3].x, d0, g0[k]); g1[k] = fmaf(win[k / 3][k % 3].y, d1, g1[k]); }
        });
      }
#pragma unroll
        for (int k = 0; k < 10; ++k) { atomicAdd(&s_dw[k][c], g0[k]); atomicAdd(&s_dw[k][c + 1], g1[k]); }
    }
    __syncthreads();
    for (int i = tid; i < 10 * a.C; i += 256) {
        const int tap = i / a.C, c = i - tap * a.C;
        const float v = s_dw[tap][c];
        if (tap < 9) atomicAdd(a.dw + tap * a.C + c, v); else atomicAdd(a.db + c, v);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        for (int c = tid; c < a.C; c += 256) {
            double g = 0.0, b = 0.0;
            for (int tt = 0; tt < kT; ++tt) { const double2 s = a.tb.bst[(size_t)tt * a.C + c]; b += s.x; g += s.y; }
            a.dgamma[c] = (float)g; a.dbeta[c] = (float)b;
        }
    }
}

// --------------------------------------------------------------------------- BatchNorm finalisation as its own launch
// Folds the replicated fp64 sums of one layer into (scale, shift) / (mean, inv_std) and applies the 4 sequential
// moving-average updates.  Running it as a separate 1-3 CTA kernel keeps device-wide fences ("last block" tickets)
// out of every producer CTA, which ncu showed as a membar stall in all of them.
struct BnFinArgs { ColMap cm; BnTables tb; int ld; BnLayer bn; double n; };
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) bn_finalize_kernel(BnFinArgs a) {
    bn_finalize(a.cm, a.tb, a.ld, a.bn.gamma, a.bn.beta, a.bn.mov_mean, a.bn.mov_var, a.n, a.bn.unbiased, a.bn.training,
                256 * gridDim.x, blockIdx.x * 256 + threadIdx.x);
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
