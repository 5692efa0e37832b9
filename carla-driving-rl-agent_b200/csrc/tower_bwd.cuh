// Backward kernels of the image tower: what tf.GradientTape computes for
// tape.gradient(loss, dynamics.trainable_variables) (core/carla_agent.py:361-365,440-444) through
// the layers of core/architectures.py:30-173.
//
// Gradient tensors hold d loss / d (activated value) in the layout of the forward tensor.  A conv's
// own BatchNorm (+ReLU6) backward is folded into its backward kernels: they rebuild
//     dR = scale * (dZ - S1/n - xhat * S2/n),  dZ = dA * [0 < z < 6],  xhat = (R - mean) * inv_std
// on load from dA, the saved raw output R and the per-(slice, channel) sums S1 = sum dZ,
// S2 = sum dZ*xhat produced by `bstat_kernel`.
#pragma once
#include "cdra_common.cuh"

namespace cdra {

// per-column BN-backward coefficients
struct BnCol { float scale, shift, mean, inv, k1, k2; };

CDRA_DEV BnCol load_bncol(const BnTables& tb, int ld, int t, int c, double inv_n) {
    BnCol r;
    const float2 a = tb.aff[(size_t)t * ld + c], b = tb.bnp[(size_t)t * ld + c];
    const double2 s = tb.bst[(size_t)t * ld + c];
    r.scale = a.x; r.shift = a.y; r.mean = b.x; r.inv = b.y;
    r.k1 = (float)(s.x * inv_n); r.k2 = (float)(s.y * inv_n);
    return r;
}
// gradient wrt the raw (pre-BN) conv output from the gradient wrt the activated value
CDRA_DEV float bn_backward_elem(float dA, float R, const BnCol& b, int clamp) {
    const float z = fmaf(R, b.scale, b.shift);
    const float dz = (!clamp || (z > 0.f && z < 6.f)) ? dA : 0.f;
    const float xhat = (R - b.mean) * b.inv;
    return b.scale * (dz - b.k1 - xhat * b.k2);
}
// `dr_kernel` applies bn_backward_elem IN PLACE to every gradient tensor right after its `bstat` launch, so the
// conv-backward kernels below read dR directly: for them the transform is the identity (the compiler drops the
// then-unused loads of R and of the BN coefficients).  Keeping the call sites documents where dR is consumed.
CDRA_DEV float make_dr(float dR, float /*R*/, const BnCol& /*b*/, int /*clamp*/) { return dR; }

// --------------------------------------------------------------------------- S1/S2 sums
template <typename T>
struct BstatArgs {
    const T* dA;          // gradient wrt activated value, same layout as R
    const T* R;           // raw tensor
    int ld, coff, C, Rt, clamp, rows_per_block;
    BnTables tb;
    unsigned* counter;    // last-block ticket (folding of the replicated sums)
};
constexpr int kBstatMaxC = 768;

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) bstat_kernel(BstatArgs<T> a) {
    CDRA_SHARED float s1[kBstatMaxC], s2[kBstatMaxC];
    const int tid = threadIdx.x, t = blockIdx.y;
    for (int i = tid; i < a.C; i += 256) { s1[i] = 0.f; s2[i] = 0.f; }
    __syncthreads();
    const int r0 = blockIdx.x * a.rows_per_block;
    const int r1 = min(a.Rt, r0 + a.rows_per_block);
    // thread <-> fixed channel pair (coalesced 2-element accesses), row lanes stride over the block's rows;
    // sums are kept in registers and folded into shared memory once per thread
    const int CP = a.C >> 1;
    const int lanes_c = CP >= 256 ? 256 : ((CP + 31) & ~31);
    const int lanes_r = 256 / lanes_c, cl = tid % lanes_c, rl = tid / lanes_c;
    for (int cp = cl; cp < CP; cp += lanes_c) {
        const int c = cp * 2;
        const float2 af0 = a.tb.aff[(size_t)t * a.ld + a.coff + c], bp0 = a.tb.bnp[(size_t)t * a.ld + a.coff + c];
        const float2 af1 = a.tb.aff[(size_t)t * a.ld + a.coff + c + 1], bp1 = a.tb.bnp[(size_t)t * a.ld + a.coff + c + 1];
        float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
        if (rl < lanes_r)
            for (int r = r0 + rl; r < r1; r += lanes_r) {
                const size_t o = ((size_t)t * a.Rt + r) * a.ld + a.coff + c;
                const float R0 = ldf(a.R + o), R1 = ldf(a.R + o + 1), d0 = ldf(a.dA + o), d1 = ldf(a.dA + o + 1);
                const float z0 = fmaf(R0, af0.x, af0.y), z1 = fmaf(R1, af1.x, af1.y);
                const float dz0 = (!a.clamp || (z0 > 0.f && z0 < 6.f)) ? d0 : 0.f;
                const float dz1 = (!a.clamp || (z1 > 0.f && z1 < 6.f)) ? d1 : 0.f;
                a0 += dz0; b0 = fmaf(dz0, (R0 - bp0.x) * bp0.y, b0);
                a1 += dz1; b1 = fmaf(dz1, (R1 - bp1.x) * bp1.y, b1);
            }
        atomicAdd(&s1[c], a0); atomicAdd(&s2[c], b0);
        atomicAdd(&s1[c + 1], a1); atomicAdd(&s2[c + 1], b1);
    }
    __syncthreads();
    for (int i = tid; i < a.C; i += 256) {
        double2* d = stat_slot(a.tb.bst, a.ld, a.counter ? stat_copy() : 0, t, a.coff + i);
        atomicAdd(&d->x, (double)s1[i]);
        atomicAdd(&d->y, (double)s2[i]);
    }
    // replicated mode (counter != null): the last block folds the replicas into replica 0, which is what the
    // backward kernels read; with counter == null everything accumulates in replica 0 directly
    if (a.counter != nullptr && last_block_ticket(a.counter, gridDim.x * gridDim.y)) {
        for (int i = tid; i < kT * a.C; i += 256) {
            const int tt = i / a.C, c = a.coff + (i - tt * a.C);
            const double2 s = stat_fold(a.tb.bst, a.ld, tt, c);
            *stat_slot(a.tb.bst, a.ld, 0, tt, c) = s;
        }
    }
}

// --------------------------------------------------------------------------- dA -> dR in place (BatchNorm + ReLU6 backward)
template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) dr_kernel(BstatArgs<T> a) {
    const int tid = threadIdx.x, t = blockIdx.y;
    const int r0 = blockIdx.x * a.rows_per_block;
    const int r1 = min(a.Rt, r0 + a.rows_per_block);
    const int CP = a.C >> 1;
    const int lanes_c = CP >= 256 ? 256 : ((CP + 31) & ~31);
    const int lanes_r = 256 / lanes_c, cl = tid % lanes_c, rl = tid / lanes_c;
    if (rl >= lanes_r) return;
    const double inv_n = 1.0 / (double)a.Rt;
    T* dA = const_cast<T*>(a.dA);
    for (int cp = cl; cp < CP; cp += lanes_c) {
        const int c = a.coff + cp * 2;
        const BnCol b0 = load_bncol(a.tb, a.ld, t, c, inv_n), b1 = load_bncol(a.tb, a.ld, t, c + 1, inv_n);
        for (int r = r0 + rl; r < r1; r += lanes_r) {
            const size_t o = ((size_t)t * a.Rt + r) * a.ld + c;
            const float2 d = ld2(dA + o), x = ld2(a.R + o);
            st2(dA + o, make_float2(bn_backward_elem(d.x, x.x, b0, a.clamp), bn_backward_elem(d.y, x.y, b1, a.clamp)));
        }
    }
}

// --------------------------------------------------------------------------- pointwise conv: data gradient
template <typename T>
struct PwBwdArgs {
    // own output (columns j -> channel colmap_c(j) of tensor `out`)
    const T* out; const T* dout; int ldo; ColMap cm; BnTables tb; int clamp;
    // input view
    ActView in; int K, Rt;
    const float* w;       // [K][N]
    T* dx; int ldx, coffx, accumulate;      // dgrad destination (may differ in ld/coff from `in`)
    float* dw; float* db; float* dgamma; float* dbeta;   // wgrad destinations
    int row_splits;
    float* partials;      // tensor-core path: per-CTA partial dW tiles (workspace scratch)
};

constexpr int kPwMaxN = 768;

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) pw_dgrad_kernel(PwBwdArgs<T> a) {
    CDRA_SHARED float As[kPwKC][kPwTM + 4];     // dR chunk  [j][row]
    CDRA_SHARED float Bs[kPwKC][kPwTN + 4];     // W^T chunk [j][k]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int t = blockIdx.y, row0 = blockIdx.x * kPwTM, k0 = blockIdx.z * kPwTN;
    const int N = a.cm.n;
    const double inv_n = 1.0 / (double)a.Rt;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 2, lj = (tid & 3) * 4;
    const int bj = tid >> 4, bk = (tid & 15) * 4;
    for (int j0 = 0; j0 < N; j0 += kPwKC) {
        const int r = row0 + lr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + lj + q;
            float v = 0.f;
            if (r < a.Rt && j < N) {
                const int c = colmap_c(a.cm, j);
                const BnCol bc = load_bncol(a.tb, a.ldo, t, c, inv_n);
                const size_t o = ((size_t)t * a.Rt + r) * a.ldo + c;
                v = make_dr(ldf(a.dout + o), ldf(a.out + o), bc, a.clamp);
            }
            As[lj + q][lr] = v;
        }
        {
            const int j = j0 + bj;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + bk + q;
                Bs[bj][bk + q] = (j < N && k < a.K) ? a.w[(size_t)k * N + colmap_w(a.cm, j)] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kPwKC; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = row0 + ty * 4 + i;
        if (r >= a.Rt) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k >= a.K) continue;
            T* d = a.dx + ((size_t)t * a.Rt + r) * a.ldx + a.coffx + k;
            stf(d, a.accumulate ? ldf(d) + acc[i][j] : acc[i][j]);
        }
    }
}

// --------------------------------------------------------------------------- pointwise conv: weight gradient
// dW[k][w(j)] += sum_rows act(in)[row][k] * dR[row][j];  the virtual row k == K is all ones -> db.
template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) pw_wgrad_kernel(PwBwdArgs<T> a) {
    CDRA_SHARED float As[kPwKC][kPwTM + 4];     // act(in) chunk [row][k]
    CDRA_SHARED float Bs[kPwKC][kPwTN + 4];     // dR chunk      [row][j]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int k0 = blockIdx.x * kPwTM, j0 = blockIdx.y * kPwTN;
    const int t = blockIdx.z / a.row_splits, sp = blockIdx.z % a.row_splits;
    const int N = a.cm.n;
    const double inv_n = 1.0 / (double)a.Rt;
    const int rows_per = (a.Rt + a.row_splits - 1) / a.row_splits;
    const int rbeg = sp * rows_per, rend = min(a.Rt, rbeg + rows_per);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 4, lc = (tid & 15) * 4;      // loader: row lr of the chunk, 4 columns from lc
    const T* in = (const T*)a.in.data;
    BnCol bc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = j0 + lc + q;
        if (j < N) bc[q] = load_bncol(a.tb, a.ldo, t, colmap_c(a.cm, j), inv_n);
    }
    for (int rr = rbeg; rr < rend; rr += kPwKC) {
        const int r = rr + lr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lc + q;
            float v = 0.f;
            if (r < rend) {
                if (k < a.K) {
                    const int c = a.in.coff + k;
                    v = act_apply(ldf(in + ((size_t)t * a.Rt + r) * a.in.ld + c),
                                  a.in.aff ? a.in.aff + (size_t)t * a.in.ld + c : nullptr, a.in.clamp);
                } else if (k == a.K) v = 1.f;
            }
            As[lr][lc + q] = v;
            const int j = j0 + lc + q;
            float d = 0.f;
            if (r < rend && j < N) {
                const size_t o = ((size_t)t * a.Rt + r) * a.ldo + colmap_c(a.cm, j);
                d = make_dr(ldf(a.dout + o), ldf(a.out + o), bc[q], a.clamp);
            }
            Bs[lr][lc + q] = d;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kPwKC; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k > a.K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jj = j0 + tx * 4 + j;
            if (jj >= N) continue;
            const int wc = colmap_w(a.cm, jj);
            if (k < a.K) atomicAdd(a.dw + (size_t)k * N + wc, acc[i][j]);
            else atomicAdd(a.db + wc, acc[i][j]);
        }
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {       // BN parameter gradients
        for (int j = tid; j < N; j += 256) {
            const int c = colmap_c(a.cm, j), wc = colmap_w(a.cm, j);
            double g = 0.0, b = 0.0;
            for (int tt = 0; tt < kT; ++tt) { const double2 s = a.tb.bst[(size_t)tt * a.ldo + c]; b += s.x; g += s.y; }
            a.dgamma[wc] = (float)g; a.dbeta[wc] = (float)b;
        }
    }
}

// --------------------------------------------------------------------------- depthwise 3x3 backward
template <typename T>
struct DwBwdArgs {
    const T* out; const T* dout; BnTables tb;      // own raw output [.. ][C] and its gradient (no ReLU after dw BN)
    ActView in;                                     // input view (C channels from in.coff)
    int B, Hi, Wi, Ho, Wo, C, stride, pad_t, pad_l;
    const float* w;                                 // [3][3][C]
    T* dx; int ldx, coffx, accumulate;
    float* dw; float* db; float* dgamma; float* dbeta;
    int ppb, ppb_w;                                 // input pixels per block (dgrad) / output pixels per block (wgrad)
};

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) dw_dgrad_kernel(DwBwdArgs<T> a) {
    const int tid = threadIdx.x, t = blockIdx.y;
    const int hiw = a.Hi * a.Wi, how = a.Ho * a.Wo, npix = a.B * hiw;
    const DwLanes L = dw_lanes(a.C, tid);
    const int p0 = blockIdx.x * a.ppb, p1 = min(npix, p0 + a.ppb);
    if (L.cl >= (a.C >> 1) || L.rl >= L.lanes_r) return;
    const int c = L.cl * 2;
    const double inv_n = 1.0 / ((double)a.B * how);
    const BnCol b0 = load_bncol(a.tb, a.C, t, c, inv_n), b1 = load_bncol(a.tb, a.C, t, c + 1, inv_n);
    float w0[9], w1[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { w0[k] = a.w[k * a.C + c]; w1[k] = a.w[k * a.C + c + 1]; }
    for (int p = p0 + L.rl; p < p1; p += L.lanes_r) {
        const int b = p / hiw, r = p - b * hiw, iy = r / a.Wi, ix = r - iy * a.Wi;
        float g0 = 0.f, g1 = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ny = iy + a.pad_t - ky;
            if (ny < 0 || ny % a.stride) continue;
            const int oy = ny / a.stride;
            if (oy >= a.Ho) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int nx = ix + a.pad_l - kx;
                if (nx < 0 || nx % a.stride) continue;
                const int ox = nx / a.stride;
                if (ox >= a.Wo) continue;
                const size_t o = (((size_t)(t * a.B + b) * a.Ho + oy) * a.Wo + ox) * a.C + c;
                const float2 dv = ld2(a.dout + o), rv = ld2(a.out + o);
                g0 = fmaf(make_dr(dv.x, rv.x, b0, 0), w0[ky * 3 + kx], g0);
                g1 = fmaf(make_dr(dv.y, rv.y, b1, 0), w1[ky * 3 + kx], g1);
            }
        }
        T* d = a.dx + ((size_t)t * npix + p) * a.ldx + a.coffx + c;
        if (a.accumulate) { const float2 old = ld2(d); g0 += old.x; g1 += old.y; }
        st2(d, make_float2(g0, g1));
    }
}

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) dw_wgrad_kernel(DwBwdArgs<T> a) {
    CDRA_SHARED float s_dw[10][kDwMaxC];        // 9 taps + bias
    const int tid = threadIdx.x, t = blockIdx.y;
    for (int i = tid; i < 10 * kDwMaxC; i += 256) (&s_dw[0][0])[i] = 0.f;
    __syncthreads();
    const int hiw = a.Hi * a.Wi, how = a.Ho * a.Wo, npix = a.B * how;
    const DwLanes L = dw_lanes(a.C, tid);
    const int p0 = blockIdx.x * a.ppb_w, p1 = min(npix, p0 + a.ppb_w);
    if (L.cl < (a.C >> 1) && L.rl < L.lanes_r) {
        const int c = L.cl * 2;
        const double inv_n = 1.0 / ((double)a.B * how);
        const BnCol b0 = load_bncol(a.tb, a.C, t, c, inv_n), b1 = load_bncol(a.tb, a.C, t, c + 1, inv_n);
        float2 f0 = make_float2(1.f, 0.f), f1 = make_float2(1.f, 0.f);
        if (a.in.aff) { f0 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c]; f1 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c + 1]; }
        float g0[10], g1[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) { g0[k] = 0.f; g1[k] = 0.f; }
        for (int p = p0 + L.rl; p < p1; p += L.lanes_r) {
            const int b = p / how, r = p - b * how, oy = r / a.Wo, ox = r - oy * a.Wo;
            const size_t o = ((size_t)t * npix + p) * a.C + c;
            const float2 dv = ld2(a.dout + o), rv = ld2(a.out + o);
            const float d0 = make_dr(dv.x, rv.x, b0, 0), d1 = make_dr(dv.y, rv.y, b1, 0);
            g0[9] += d0; g1[9] += d1;
            const T* base = (const T*)a.in.data + ((size_t)(t * a.B + b) * hiw) * a.in.ld + a.in.coff + c;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = oy * a.stride - a.pad_t + ky;
                if (iy < 0 || iy >= a.Hi) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = ox * a.stride - a.pad_l + kx;
                    if (ix < 0 || ix >= a.Wi) continue;
                    float2 v = ld2(base + ((size_t)iy * a.Wi + ix) * a.in.ld);
                    v.x = fmaf(v.x, f0.x, f0.y); v.y = fmaf(v.y, f1.x, f1.y);
                    if (a.in.clamp) { v.x = relu6f(v.x); v.y = relu6f(v.y); }
                    g0[ky * 3 + kx] = fmaf(v.x, d0, g0[ky * 3 + kx]);
                    g1[ky * 3 + kx] = fmaf(v.y, d1, g1[ky * 3 + kx]);
                }
            }
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

// --------------------------------------------------------------------------- pass-through half (stride-1 units)
template <typename T>
struct PassBwdArgs {
    const T* dout; int ldo;     // gradient of the unit output
    T* dx; int ldx;             // gradient of the unit input (left half written)
    int half, Rt;
};
template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) pass_bwd_kernel(PassBwdArgs<T> a) {
    const int t = blockIdx.y, QP = a.half >> 1;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)a.Rt * QP) return;
    const int qp = (int)(idx % QP);
    const long long r = idx / QP;
    const T* s = a.dout + ((size_t)t * a.Rt + r) * a.ldo;
    T* d = a.dx + ((size_t)t * a.Rt + r) * a.ldx + 2 * qp;
    d[0] = s[qp];
    d[1] = s[a.half + qp];
}

// --------------------------------------------------------------------------- maxpool backward (first max wins)
template <typename T>
struct PoolBwdArgs {
    ActView in;                 // stem raw (+affine, ReLU6)
    const T* dpool;             // [kT*B*Ho*Wo][C]
    T* dstem;                   // [kT*B*Hi*Wi][C] gradient wrt the activated stem output
    int B, Hi, Wi, Ho, Wo, C, pad_t, pad_l;
};
template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) pool_bwd_kernel(PoolBwdArgs<T> a) {
    const int t = blockIdx.y;
    const int hiw = a.Hi * a.Wi;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)a.B * hiw * a.C) return;
    const int c = (int)(idx % a.C);
    const long long p = idx / a.C;
    const int b = (int)(p / hiw), r = (int)(p - (long long)b * hiw), y = r / a.Wi, x = r - y * a.Wi;
    const T* base = (const T*)a.in.data + ((size_t)(t * a.B + b) * hiw) * a.in.ld + a.in.coff + c;
    const float2* af = a.in.aff ? a.in.aff + (size_t)t * a.in.ld + a.in.coff + c : nullptr;
    float g = 0.f;
    const int oy_lo = max(0, (y + a.pad_t - 1) / 2), oy_hi = min(a.Ho - 1, (y + a.pad_t) / 2);
    const int ox_lo = max(0, (x + a.pad_l - 1) / 2), ox_hi = min(a.Wo - 1, (x + a.pad_l) / 2);
    for (int oy = oy_lo; oy <= oy_hi; ++oy)
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
            float best = -INFINITY; int by = -1, bx = -1;
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = oy * 2 - a.pad_t + ky;
                if (iy < 0 || iy >= a.Hi) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = ox * 2 - a.pad_l + kx;
                    if (ix < 0 || ix >= a.Wi) continue;
                    const float v = act_apply(ldf(base + ((size_t)iy * a.Wi + ix) * a.in.ld), af, a.in.clamp);
                    if (v > best) { best = v; by = iy; bx = ix; }
                }
            }
            if (by == y && bx == x)
                g += ldf(a.dpool + (((size_t)(t * a.B + b) * a.Ho + oy) * a.Wo + ox) * a.C + c);
        }
    stf(a.dstem + ((size_t)t * a.B * hiw + p) * a.C + c, g);
}

// --------------------------------------------------------------------------- stem weight gradient
template <typename T, typename TIn>
struct StemBwdArgs {
    const TIn* img; int B, H, W, Ho, Wo;
    const T* out; const T* dout; BnTables tb;       // stem raw + gradient wrt activated
    float* dw; float* db; float* dgamma; float* dbeta;
    int pix_per_block;
};
constexpr int kStemWgThreads = 672;     // >= 27*24 = 648
constexpr int kStemWgP = 128;

template <typename T, typename TIn>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(kStemWgThreads) stem_wgrad_kernel(StemBwdArgs<T, TIn> a) {
    CDRA_SHARED float s_in[kStemWgP][28];
    CDRA_SHARED float s_dr[kStemWgP][kStemC];
    CDRA_SHARED float s_lut[256];
    const int tid = threadIdx.x, t = blockIdx.y;
    for (int i = tid; i < 256; i += kStemWgThreads) s_lut[i] = __fdiv_rn((float)i, 255.f);
    const int hw = a.Ho * a.Wo, Rt = a.B * hw;
    const double inv_n = 1.0 / (double)Rt;
    const int p0 = blockIdx.x * a.pix_per_block, p1 = min(Rt, p0 + a.pix_per_block);
    const int k = tid / kStemC, n = tid % kStemC;       // thread owns dW[k][n] (k < 27) ; k == 27 -> db[n]
    float acc = 0.f;
    __syncthreads();
    for (int pp = p0; pp < p1; pp += kStemWgP) {
        for (int i = tid; i < kStemWgP * 27; i += kStemWgThreads) {
            const int pl = i / 27, kk = i - pl * 27, p = pp + pl;
            float v = 0.f;
            if (p < p1) {
                const int b = p / hw, r = p - b * hw, oy = r / a.Wo, ox = r - oy * a.Wo;
                const int ky = kk / 9, rem = kk - ky * 9;
                v = img_to_float(a.img[((size_t)(b * kT + t) * a.H + 2 * oy + ky) * a.W * 3 + (size_t)2 * ox * 3 + rem], s_lut);
            }
            s_in[pl][kk] = v;
        }
        for (int i = tid; i < kStemWgP * kStemC; i += kStemWgThreads) {
            const int pl = i / kStemC, c = i - pl * kStemC, p = pp + pl;
            float v = 0.f;
            if (p < p1) {
                const BnCol bc = load_bncol(a.tb, kStemC, t, c, inv_n);
                const size_t o = ((size_t)t * Rt + p) * kStemC + c;
                v = make_dr(ldf(a.dout + o), ldf(a.out + o), bc, 1);
            }
            s_dr[pl][c] = v;
        }
        __syncthreads();
        if (k < 27) {
#pragma unroll 8
            for (int pl = 0; pl < kStemWgP; ++pl) acc = fmaf(s_in[pl][k], s_dr[pl][n], acc);
        } else if (k == 27) {
#pragma unroll 8
            for (int pl = 0; pl < kStemWgP; ++pl) acc += s_dr[pl][n];
        }
        __syncthreads();
    }
    if (k < 27) atomicAdd(a.dw + k * kStemC + n, acc);
    else if (k == 27) atomicAdd(a.db + n, acc);
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid < kStemC) {
        double g = 0.0, b = 0.0;
        for (int tt = 0; tt < kT; ++tt) { const double2 s = a.tb.bst[(size_t)tt * kStemC + tid]; b += s.x; g += s.y; }
        a.dgamma[tid] = (float)g; a.dbeta[tid] = (float)b;
    }
}

// --------------------------------------------------------------------------- global-average-pool backward
template <typename T>
struct GapBwdArgs { const float* dgap; T* dhead; int B, HW, C; };
template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) gap_bwd_kernel(GapBwdArgs<T> a) {
    const int t = blockIdx.y;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)a.B * a.HW * a.C) return;
    const int c = (int)(idx % a.C);
    const long long fp = idx / a.C;
    const int b = (int)(fp / a.HW);
    stf(a.dhead + ((size_t)t * a.B * a.HW + fp) * a.C + c, a.dgap[((size_t)t * a.B + b) * a.C + c] / (float)a.HW);
}

}  // namespace cdra
