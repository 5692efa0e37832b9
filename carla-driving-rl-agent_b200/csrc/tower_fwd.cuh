// Forward kernels of the time-shared ShuffleNet-v2 image tower
// (restates core/architectures.py:30-173 of the reference as sm_100a kernels).
//
// Layout: activations NHWC, frames ordered [slice t][sample b] so that the per-call BatchNorm
// statistics of the reference (one BN call per time slice, core/architectures.py:44-57) are
// reductions over contiguous row ranges.  Every conv stores its *raw* (pre-BN) output once; the
// per-(slice, channel) BN affine (+ReLU6) is applied by the consumer when it loads the tensor.
// Statistics are accumulated in fp64 by the producer's epilogue and finalised by its last block.
#pragma once
#include "cdra_common.cuh"

namespace cdra {

struct BnLayer {          // parameters of the BatchNorm that follows a conv
    const float* gamma;
    const float* beta;
    float* mov_mean;
    float* mov_var;
    unsigned* counter;    // last-block ticket
    int training;
    int unbiased;
};

// --------------------------------------------------------------------------- stem: 3x3 s2 VALID, 3 -> 24
template <typename T, typename TIn>
struct StemArgs {
    const TIn* img;       // [B][kT][H][W][3] (reference layout, core/networks.py:237-245)
    int B, H, W, Ho, Wo;
    const float* w;       // [3][3][3][24]
    const float* bias;    // [24]
    T* out;               // [kT*B*Ho*Wo][24]
    BnTables tb;
    BnLayer bn;
};

CDRA_DEV float img_to_float(uint8_t v, const float* lut) { return lut[v]; }
CDRA_DEV float img_to_float(float v, const float*) { return v; }

constexpr int kStemC = 24;
constexpr int kStemPPT = 4;     // pixels per thread

template <typename T, typename TIn>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(128) stem_fwd_kernel(StemArgs<T, TIn> a) {
    CDRA_SHARED float s_w[27 * kStemC];
    CDRA_SHARED float s_b[kStemC];
    CDRA_SHARED float s_lut[256];
    CDRA_SHARED float s_sum[kStemC], s_sq[kStemC];
    const int tid = threadIdx.x;
    for (int i = tid; i < 27 * kStemC; i += 128) s_w[i] = a.w[i];
    if (tid < kStemC) { s_b[tid] = a.bias[tid]; s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
    for (int i = tid; i < 256; i += 128) s_lut[i] = __fdiv_rn((float)i, 255.f);   // reference feeds u8/255
    __syncthreads();

    const int t = blockIdx.y;
    const int hw = a.Ho * a.Wo;
    const int Rt = a.B * hw;
    float lsum[kStemC], lsq[kStemC];
#pragma unroll
    for (int c = 0; c < kStemC; ++c) { lsum[c] = 0.f; lsq[c] = 0.f; }

    for (int i = 0; i < kStemPPT; ++i) {
        const int p = (blockIdx.x * kStemPPT + i) * 128 + tid;
        if (p < Rt) {
            const int b = p / hw, r = p - b * hw, oy = r / a.Wo, ox = r - oy * a.Wo;
            const TIn* src = a.img + ((size_t)(b * kT + t) * a.H + 2 * oy) * a.W * 3 + (size_t)2 * ox * 3;
            float acc[kStemC];
#pragma unroll
            for (int c = 0; c < kStemC; ++c) acc[c] = s_b[c];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const TIn* row = src + (size_t)ky * a.W * 3;
#pragma unroll
                for (int kx = 0; kx < 9; ++kx) {      // 3 pixels x 3 channels are contiguous
                    const float x = img_to_float(row[kx], s_lut);
                    const float* wp = s_w + (ky * 9 + kx) * kStemC;
#pragma unroll
                    for (int c = 0; c < kStemC; ++c) acc[c] = fmaf(x, wp[c], acc[c]);
                }
            }
            T* dst = a.out + ((size_t)t * Rt + p) * kStemC;
#pragma unroll
            for (int c = 0; c < kStemC; ++c) {
                const float v = rnd(acc[c], dst);
                stf(dst + c, acc[c]);
                lsum[c] += v; lsq[c] += v * v;
            }
        }
    }
    if (!a.bn.training) return;        // inference: moving statistics, no batch sums (block-uniform)
#pragma unroll
    for (int c = 0; c < kStemC; ++c) {
        const float s = warp_sum(lsum[c]), q = warp_sum(lsq[c]);
        if ((tid & 31) == 0) { atomicAdd(&s_sum[c], s); atomicAdd(&s_sq[c], q); }
    }
    __syncthreads();
    if (tid < kStemC) {
        double2* dst = stat_slot(a.tb.fst, kStemC, stat_copy(), t, tid);
        atomicAdd(&dst->x, (double)s_sum[tid]);
        atomicAdd(&dst->y, (double)s_sq[tid]);
    }
    const unsigned total = gridDim.x * gridDim.y;
    if (a.bn.counter != nullptr && last_block_ticket(a.bn.counter, total)) {
        ColMap cm{kStemC, 0, 0, 0};
        bn_finalize(cm, a.tb, kStemC, a.bn.gamma, a.bn.beta, a.bn.mov_mean, a.bn.mov_var, (double)Rt,
                    a.bn.unbiased, a.bn.training, 128, tid);
    }
}

// --------------------------------------------------------------------------- maxpool 3x3 s2 SAME
template <typename T>
struct PoolArgs {
    ActView in;           // stem raw + BN affine + ReLU6
    int B, Hi, Wi, Ho, Wo, C, pad_t, pad_l;
    T* out;               // activated values, [kT*B*Ho*Wo][C]
};

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) pool_fwd_kernel(PoolArgs<T> a) {
    const int t = blockIdx.y;
    const int CP = a.C >> 1;
    const int how = a.Ho * a.Wo;
    const long long total = (long long)a.B * how * CP;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int cp = (int)(idx % CP);
    const long long p = idx / CP;
    const int b = (int)(p / how), r = (int)(p - (long long)b * how), oy = r / a.Wo, ox = r - oy * a.Wo;
    const int c = cp * 2;
    const T* base = (const T*)a.in.data + ((size_t)(t * a.B + b) * a.Hi * a.Wi) * a.in.ld + a.in.coff + c;
    const float2* af = a.in.aff ? a.in.aff + (size_t)t * a.in.ld + a.in.coff + c : nullptr;
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - a.pad_t + ky;
        if (iy < 0 || iy >= a.Hi) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - a.pad_l + kx;
            if (ix < 0 || ix >= a.Wi) continue;
            const T* q = base + ((size_t)iy * a.Wi + ix) * a.in.ld;
            m0 = fmaxf(m0, act_apply(ldf(q), af, a.in.clamp));
            m1 = fmaxf(m1, act_apply(ldf(q + 1), af ? af + 1 : nullptr, a.in.clamp));
        }
    }
    T* dst = a.out + ((size_t)t * a.B * how + p) * a.C + c;
    stf(dst, m0); stf(dst + 1, m1);
}

// --------------------------------------------------------------------------- pointwise (1x1) conv, CUDA-core path
// out_raw[row][colmap_c(j)] = sum_k act(in[row][k]) * W[k][colmap_w(j)] + bias[colmap_w(j)]
template <typename T>
struct PwArgs {
    ActView in;           // K channels
    int K, Rt;            // rows per slice
    const float* w;       // [K][N]
    const float* bias;    // [N]
    ColMap cm;            // N output columns -> destination channels
    T* out;               // destination tensor, ldo channels per row
    int ldo;
    BnTables tb;          // tables of the destination tensor (indexed by destination channel)
    BnLayer bn;
    int do_stats;
};

constexpr int kPwTM = 64, kPwTN = 64, kPwKC = 16;

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) pw_fwd_kernel(PwArgs<T> a) {
    CDRA_SHARED float As[kPwKC][kPwTM + 4];
    CDRA_SHARED float Bs[kPwKC][kPwTN + 4];
    CDRA_SHARED float s_sum[kPwTN], s_sq[kPwTN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int t = blockIdx.y, row0 = blockIdx.x * kPwTM, col0 = blockIdx.z * kPwTN;
    const int N = a.cm.n;
    if (tid < kPwTN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const T* in = (const T*)a.in.data;
    const int lr = tid >> 2, lk = (tid & 3) * 4;        // A loader: row lr, 4 k's from lk
    const int br = tid >> 4, bc = (tid & 15) * 4;       // B loader: k br, 4 cols from bc
    for (int k0 = 0; k0 < a.K; k0 += kPwKC) {
        {
            const int r = row0 + lr;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + lk + q;
                float v = 0.f;
                if (r < a.Rt && k < a.K) {
                    const int c = a.in.coff + k;
                    const float raw = ldf(in + ((size_t)t * a.Rt + r) * a.in.ld + c);
                    v = act_apply(raw, a.in.aff ? a.in.aff + (size_t)t * a.in.ld + c : nullptr, a.in.clamp);
                }
                As[lk + q][lr] = v;
            }
            const int k = k0 + br;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = col0 + bc + q;
                Bs[br][bc + q] = (k < a.K && j < N) ? a.w[(size_t)k * N + colmap_w(a.cm, j)] : 0.f;
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
    // epilogue: bias, store raw, statistics over the stored values
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = col0 + tx * 4 + j;
        if (col >= N) continue;
        const int c = colmap_c(a.cm, col);
        const float bj = a.bias[colmap_w(a.cm, col)];
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = row0 + ty * 4 + i;
            if (r < a.Rt) {
                T* dst = a.out + ((size_t)t * a.Rt + r) * a.ldo + c;
                const float v = acc[i][j] + bj;
                stf(dst, v);
                const float vr = rnd(v, dst);
                s += vr; q += vr * vr;
            }
        }
        if (a.do_stats) { atomicAdd(&s_sum[tx * 4 + j], s); atomicAdd(&s_sq[tx * 4 + j], q); }
    }
    if (!a.do_stats) return;
    __syncthreads();
    if (tid < kPwTN && col0 + tid < N) {
        double2* dst = stat_slot(a.tb.fst, a.ldo, stat_copy(), t, colmap_c(a.cm, col0 + tid));
        atomicAdd(&dst->x, (double)s_sum[tid]);
        atomicAdd(&dst->y, (double)s_sq[tid]);
    }
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    if (a.bn.counter != nullptr && last_block_ticket(a.bn.counter, total))
        bn_finalize(a.cm, a.tb, a.ldo, a.bn.gamma, a.bn.beta, a.bn.mov_mean, a.bn.mov_var, (double)a.Rt,
                    a.bn.unbiased, a.bn.training, 256, tid);
}

// --------------------------------------------------------------------------- depthwise 3x3, stride 1|2, TF SAME
template <typename T>
struct DwArgs {
    ActView in;
    int B, Hi, Wi, Ho, Wo, C, stride, pad_t, pad_l;
    const float* w;       // [3][3][C]
    const float* bias;    // [C]
    T* out;               // [kT*B*Ho*Wo][C] raw
    BnTables tb;
    BnLayer bn;
    int ppb;              // output pixels per block
};

constexpr int kDwItems = 4;     // legacy constant (grid sizing of the weight-gradient kernel)
constexpr int kDwMaxC = 256;

// thread <-> fixed channel pair (weights / affine in registers, coalesced 2-channel accesses), row lanes loop
// over the block's output pixels; statistics accumulate in registers.
struct DwLanes { int lanes_c, lanes_r, cl, rl; };
CDRA_DEV DwLanes dw_lanes(int C, int tid) {
    DwLanes l;
    const int CP = C >> 1;
    l.lanes_c = (CP + 31) & ~31;
    if (l.lanes_c > 256) l.lanes_c = 256;
    l.lanes_r = 256 / l.lanes_c; l.cl = tid % l.lanes_c; l.rl = tid / l.lanes_c;
    return l;
}

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) dw_fwd_kernel(DwArgs<T> a) {
    CDRA_SHARED float s_sum[kDwMaxC], s_sq[kDwMaxC];
    const int tid = threadIdx.x, t = blockIdx.y;
    for (int i = tid; i < a.C; i += 256) { s_sum[i] = 0.f; s_sq[i] = 0.f; }
    __syncthreads();
    const int how = a.Ho * a.Wo, npix = a.B * how;
    const DwLanes L = dw_lanes(a.C, tid);
    const int p0 = blockIdx.x * a.ppb, p1 = min(npix, p0 + a.ppb);
    if (L.cl < (a.C >> 1) && L.rl < L.lanes_r) {
        const int c = L.cl * 2;
        float w0[9], w1[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { w0[k] = a.w[k * a.C + c]; w1[k] = a.w[k * a.C + c + 1]; }
        const float bias0 = a.bias[c], bias1 = a.bias[c + 1];
        float2 f0 = make_float2(1.f, 0.f), f1 = make_float2(1.f, 0.f);
        if (a.in.aff) { f0 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c]; f1 = a.in.aff[(size_t)t * a.in.ld + a.in.coff + c + 1]; }
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
        for (int p = p0 + L.rl; p < p1; p += L.lanes_r) {
            const int b = p / how, r = p - b * how, oy = r / a.Wo, ox = r - oy * a.Wo;
            const T* base = (const T*)a.in.data + ((size_t)(t * a.B + b) * a.Hi * a.Wi) * a.in.ld + a.in.coff + c;
            float a0 = bias0, a1 = bias1;
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
                    a0 = fmaf(v.x, w0[ky * 3 + kx], a0);
                    a1 = fmaf(v.y, w1[ky * 3 + kx], a1);
                }
            }
            T* dst = a.out + ((size_t)t * npix + p) * a.C + c;
            st2(dst, make_float2(a0, a1));
            const float v0 = rnd(a0, dst), v1 = rnd(a1, dst);
            s0 += v0; q0 = fmaf(v0, v0, q0); s1 += v1; q1 = fmaf(v1, v1, q1);
        }
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
        bn_finalize(cm, a.tb, a.C, a.bn.gamma, a.bn.beta, a.bn.mov_mean, a.bn.mov_var, (double)npix,
                    a.bn.unbiased, a.bn.training, 256, tid);
    }
}

// --------------------------------------------------------------------------- stride-1 unit pass-through half
// concat = [shortcut | branch], shuffle: out[(q%2)*C/2 + q/2] = concat[q]; the shortcut half is a pure
// index permutation of the unit input's left half (bit-exact copy of the raw values + their BN tables).
template <typename T>
struct PassArgs {
    const T* in;  int ldi;        // unit input tensor, left half = channels [0, half)
    T* out;       int ldo;        // unit output tensor (C = 2*half channels)
    int half, Rt;
    const float2* aff_in; float2* aff_out;     // [kT][ld] tables (may be null when the input is plain)
    const float2* bnp_in; float2* bnp_out;
};

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) pass_fwd_kernel(PassArgs<T> a) {
    const int t = blockIdx.y;
    const int QP = a.half >> 1;
    const long long total = (long long)a.Rt * QP;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.x == 0) {
        for (int q = threadIdx.x; q < a.half; q += 256) {
            const int d = (q & 1) * a.half + (q >> 1);
            a.aff_out[(size_t)t * a.ldo + d] = a.aff_in ? a.aff_in[(size_t)t * a.ldi + q] : make_float2(1.f, 0.f);
            a.bnp_out[(size_t)t * a.ldo + d] = a.bnp_in ? a.bnp_in[(size_t)t * a.ldi + q] : make_float2(0.f, 1.f);
        }
    }
    if (idx >= total) return;
    const int qp = (int)(idx % QP);
    const long long r = idx / QP;
    const T* src = a.in + ((size_t)t * a.Rt + r) * a.ldi + 2 * qp;
    T* dst = a.out + ((size_t)t * a.Rt + r) * a.ldo;
    dst[qp] = src[0];                    // even q -> left half position q/2
    dst[a.half + qp] = src[1];           // odd q  -> right half position q/2
}

// --------------------------------------------------------------------------- global average pool
template <typename T>
struct GapArgs {
    ActView in;           // head conv raw + affine + ReLU6, [kT*B][HW][C]
    int B, HW, C;
    float* out;           // [kT*B][C] fp32
};

template <typename T>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) gap_fwd_kernel(GapArgs<T> a) {
    const int t = blockIdx.y;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)a.B * a.C) return;
    const int c = (int)(idx % a.C), b = (int)(idx / a.C);
    const T* base = (const T*)a.in.data + ((size_t)(t * a.B + b) * a.HW) * a.in.ld + a.in.coff + c;
    const float2* af = a.in.aff ? a.in.aff + (size_t)t * a.in.ld + a.in.coff + c : nullptr;
    float s = 0.f;
    for (int p = 0; p < a.HW; ++p) s += act_apply(ldf(base + (size_t)p * a.in.ld), af, a.in.clamp);
    a.out[((size_t)t * a.B + b) * a.C + c] = s / (float)a.HW;
}

}  // namespace cdra
