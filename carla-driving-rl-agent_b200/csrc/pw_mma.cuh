// bf16 tensor-core versions of the three pointwise-conv (1x1) kernels: forward, data gradient, weight
// gradient (core/architectures.py:130,134,140,170 and their tape gradients).  Legacy-path MMA
// (mma.sync.m16n8k16, fp32 accumulate) with register-staged, software-pipelined operand loads so that
// the BatchNorm affine / ReLU6 / BN-backward transforms are applied on the way into shared memory.
// Only compiled for the device (no emulator counterpart): the CUDA-core kernels in tower_fwd.cuh /
// tower_bwd.cuh are the logic reference, and GPU tests compare both against the oracle.
#pragma once
#ifndef CDRA_EMU
#include "tower_fwd.cuh"
#include "tower_bwd.cuh"

namespace cdra {

CDRA_DEV uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
CDRA_DEV float2 unpack_bf16(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
CDRA_DEV void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
CDRA_DEV void ldsm_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
CDRA_DEV void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// ---- bf16 copies of the pointwise weights, rebuilt once per forward (weights change every SGD step)
//   wt [Np8][Kp32] : wt[j][k] = W[k][colmap_w(j)]   (B operand of the forward GEMM, k contiguous)
//   wn [Kp8][Np32] : wn[k][j] = W[k][colmap_w(j)]   (B operand of the data-gradient GEMM, j contiguous)
struct WPrepLayer { const float* w; bf16* wt; bf16* wn; int K, N, Kp, Np, split; };
constexpr int kWPrepMax = 40;
struct WPrepArgs { WPrepLayer l[kWPrepMax]; int n; };

CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) wprep_kernel(WPrepArgs a) {
    const WPrepLayer L = a.l[blockIdx.x];
    ColMap cm{L.N, L.split, 0, 0};
    const int Np8 = (L.N + 7) & ~7, Kp8 = (L.K + 7) & ~7;
    for (int i = blockIdx.y * 256 + threadIdx.x; i < Np8 * L.Kp; i += gridDim.y * 256) {
        const int j = i / L.Kp, k = i - j * L.Kp;
        L.wt[i] = __float2bfloat16_rn((j < L.N && k < L.K) ? L.w[(size_t)k * L.N + colmap_w(cm, j)] : 0.f);
    }
    for (int i = blockIdx.y * 256 + threadIdx.x; i < Kp8 * L.Np; i += gridDim.y * 256) {
        const int k = i / L.Np, j = i - k * L.Np;
        L.wn[i] = __float2bfloat16_rn((j < L.N && k < L.K) ? L.w[(size_t)k * L.N + colmap_w(cm, j)] : 0.f);
    }
}

// tile: 128 rows x TN output columns per CTA (TN = 64 for narrow layers keeps registers / smem small so that more
// CTAs are resident), reduction chunks of 32, 8 warps x 16 rows
constexpr int kMmTM = 128, kMmKC = 32, kMmLd = kMmKC + 8;

// ======================================================================================== forward
struct PwMmaFwdArgs {
    PwArgs<bf16> a;
    const bf16* wt; int Kp;
};

template <int TN>
CDRA_KERNEL __launch_bounds__(256, TN == 64 ? 3 : 2) pw_fwd_mma_kernel(PwMmaFwdArgs pa) {
    constexpr int NB = TN / 8, kCLd = TN + 8;
    constexpr int kAB = 2 * kMmTM * kMmLd + 2 * TN * kMmLd, kC = kMmTM * kCLd;
    const PwArgs<bf16>& a = pa.a;
    __shared__ __align__(16) bf16 smem[kAB > kC ? kAB : kC];     // As[2] | Bs[2]; reused as the output tile Cs
    __shared__ float s_scale[480], s_shift[480];
    __shared__ float s_sum[TN], s_sq[TN];
    bf16* As = smem;
    bf16* Bs = smem + 2 * kMmTM * kMmLd;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.y, row0 = blockIdx.x * kMmTM, col0 = blockIdx.z * TN;
    const int N = a.cm.n, K = a.K;
    const int ncols = min(TN, N - col0), nblk = (ncols + 7) >> 3;
    for (int k = tid; k < pa.Kp; k += 256) {
        float sc = 1.f, sh = 0.f;
        if (a.in.aff && k < K) { const float2 f = a.in.aff[(size_t)t * a.in.ld + a.in.coff + k]; sc = f.x; sh = f.y; }
        s_scale[k] = sc; s_shift[k] = sh;
    }
    if (tid < TN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
    __syncthreads();

    float acc[NB][4];
#pragma unroll
    for (int i = 0; i < NB; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

    const int p = tid & 15, rb = tid >> 4;               // loader: k pair p, rows rb + 16 i
    const bf16* in = (const bf16*)a.in.data;
    const int clampf = a.in.clamp;
    uint32_t ra[8], rbw[TN / 16];
    auto load_chunk = [&](int k0) {
        const int k = k0 + 2 * p;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = row0 + rb + 16 * i;
            uint32_t v = 0u;
            if (r < a.Rt && k < K) {
                const uint32_t raw = *reinterpret_cast<const uint32_t*>(in + ((size_t)t * a.Rt + r) * a.in.ld + a.in.coff + k);
                float2 f = unpack_bf16(raw);
                f.x = fmaf(f.x, s_scale[k], s_shift[k]); f.y = fmaf(f.y, s_scale[k + 1], s_shift[k + 1]);
                if (clampf) { f.x = relu6f(f.x); f.y = relu6f(f.y); }
                v = pack_bf16(f.x, f.y);
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < TN / 16; ++i) {
            const int n = col0 + rb + 16 * i;
            rbw[i] = (rb + 16 * i < nblk * 8 && k < pa.Kp) ? *reinterpret_cast<const uint32_t*>(pa.wt + (size_t)n * pa.Kp + k) : 0u;
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint32_t*>(As + (size_t)buf * kMmTM * kMmLd + (rb + 16 * i) * kMmLd + 2 * p) = ra[i];
#pragma unroll
        for (int i = 0; i < TN / 16; ++i)
            *reinterpret_cast<uint32_t*>(Bs + (size_t)buf * TN * kMmLd + (rb + 16 * i) * kMmLd + 2 * p) = rbw[i];
    };
    const int nchunks = (K + kMmKC - 1) / kMmKC;
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int kc = 0; kc < nchunks; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nchunks) load_chunk((kc + 1) * kMmKC);
        const bf16* Ab = As + (size_t)buf * kMmTM * kMmLd;
        const bf16* Bb = Bs + (size_t)buf * TN * kMmLd;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t af[4];
            ldsm_x4(af, Ab + (warp * 16 + (lane & 15)) * kMmLd + ks * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int nb2 = 0; nb2 < NB / 2; ++nb2) {
                if (nb2 * 2 < nblk) {
                    uint32_t bf[4];
                    const int mi = lane >> 3;
                    ldsm_x4(bf, Bb + (nb2 * 16 + (mi >> 1) * 8 + (lane & 7)) * kMmLd + ks * 16 + (mi & 1) * 8);
                    mma_bf16(acc[nb2 * 2], af, bf[0], bf[1]);
                    mma_bf16(acc[nb2 * 2 + 1], af, bf[2], bf[3]);
                }
            }
        }
        if (kc + 1 < nchunks) store_chunk(buf ^ 1);
        __syncthreads();
    }
    // ---- epilogue: bias, stage the bf16 tile in shared memory, column statistics, coalesced store
    bf16* Cs = smem;
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        if (nb < nblk) {
            const int j = nb * 8 + 2 * tg;
            const float b0 = (col0 + j < N) ? a.bias[colmap_w(a.cm, col0 + j)] : 0.f;
            const float b1 = (col0 + j + 1 < N) ? a.bias[colmap_w(a.cm, col0 + j + 1)] : 0.f;
            *reinterpret_cast<uint32_t*>(Cs + (warp * 16 + g) * kCLd + j) = pack_bf16(acc[nb][0] + b0, acc[nb][1] + b1);
            *reinterpret_cast<uint32_t*>(Cs + (warp * 16 + g + 8) * kCLd + j) = pack_bf16(acc[nb][2] + b0, acc[nb][3] + b1);
        }
    }
    __syncthreads();
    const int rows = min(kMmTM, a.Rt - row0);
    {   // one pass over the staged tile: thread <-> fixed column pair (destination channels resolved once), row lanes
        // stride the rows; column statistics (over the stored, rounded values) accumulate in registers on the way out
        constexpr int kPairs = TN / 2, kRowLanes = 256 / kPairs;
        const int pj = tid % kPairs, rl = tid / kPairs, j = pj * 2;
        if (j < ncols) {
            const bool two = j + 1 < ncols;
            const int c0 = colmap_c(a.cm, col0 + j), c1 = two ? colmap_c(a.cm, col0 + j + 1) : 0;
            const bool vec = two && c1 == c0 + 1 && !(c0 & 1);
            bf16* orow = a.out + ((size_t)t * a.Rt + row0) * a.ldo;
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
            for (int r = rl; r < rows; r += kRowLanes) {
                const uint32_t u = *reinterpret_cast<const uint32_t*>(Cs + r * kCLd + j);
                const float2 v = unpack_bf16(u);
                bf16* row = orow + (size_t)r * a.ldo;
                if (vec) *reinterpret_cast<uint32_t*>(row + c0) = u;
                else { row[c0] = Cs[r * kCLd + j]; if (two) row[c1] = Cs[r * kCLd + j + 1]; }
                s0 += v.x; q0 = fmaf(v.x, v.x, q0); s1 += v.y; q1 = fmaf(v.y, v.y, q1);
            }
            if (a.do_stats) {
                atomicAdd(&s_sum[j], s0); atomicAdd(&s_sq[j], q0);
                if (two) { atomicAdd(&s_sum[j + 1], s1); atomicAdd(&s_sq[j + 1], q1); }
            }
        }
    }
    if (!a.do_stats) return;           // inference (block-uniform)
    __syncthreads();
    if (tid < ncols) {
        double2* dst = stat_slot(a.tb.fst, a.ldo, stat_copy(), t, colmap_c(a.cm, col0 + tid));
        atomicAdd(&dst->x, (double)s_sum[tid]);
        atomicAdd(&dst->y, (double)s_sq[tid]);
    }
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    if (a.bn.counter != nullptr && last_block_ticket(a.bn.counter, total))
        bn_finalize(a.cm, a.tb, a.ldo, a.bn.gamma, a.bn.beta, a.bn.mov_mean, a.bn.mov_var, (double)a.Rt,
                    a.bn.unbiased, a.bn.training, 256, tid);
}

// ======================================================================================== data gradient
struct PwMmaBwdArgs {
    PwBwdArgs<bf16> a;
    const bf16* wn; int Np;          // [Kp8][Np] bf16
};

// dR for output column pair (j, j+1) of row `grow` (global row index), as packed bf16x2
CDRA_DEV uint32_t load_dr_pair(const PwBwdArgs<bf16>& a, size_t grow, int j, int N, const BnCol& c0, const BnCol& c1, int ch0, int ch1) {
    float d0 = 0.f, d1 = 0.f;
    const size_t o = grow * a.ldo;
    if (j + 1 < N && ch1 == ch0 + 1 && !(ch0 & 1)) {
        const float2 dv = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.dout + o + ch0));
        const float2 rv = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.out + o + ch0));
        d0 = make_dr(dv.x, rv.x, c0, a.clamp); d1 = make_dr(dv.y, rv.y, c1, a.clamp);
    } else {
        if (j < N) d0 = make_dr(__bfloat162float(a.dout[o + ch0]), __bfloat162float(a.out[o + ch0]), c0, a.clamp);
        if (j + 1 < N) d1 = make_dr(__bfloat162float(a.dout[o + ch1]), __bfloat162float(a.out[o + ch1]), c1, a.clamp);
    }
    return pack_bf16(d0, d1);
}

template <int TK>
CDRA_KERNEL __launch_bounds__(256, TK == 64 ? 3 : 2) pw_dgrad_mma_kernel(PwMmaBwdArgs pa) {
    constexpr int NB = TK / 8, kCLd = TK + 8;
    constexpr int kAB = 2 * kMmTM * kMmLd + 2 * TK * kMmLd, kC = kMmTM * kCLd;
    const PwBwdArgs<bf16>& a = pa.a;
    __shared__ __align__(16) bf16 smem[kAB > kC ? kAB : kC];
    bf16* As = smem;                                   // dR chunk   [row][n]
    bf16* Bs = smem + 2 * kMmTM * kMmLd;               // W chunk    [k][n]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.y, row0 = blockIdx.x * kMmTM, k0t = blockIdx.z * TK;
    const int N = a.cm.n, K = a.K;
    const int kcols = min(TK, K - k0t), kblk = (kcols + 7) >> 3;
    const double inv_n = 1.0 / (double)a.Rt;
    float acc[NB][4];
#pragma unroll
    for (int i = 0; i < NB; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    const int p = tid & 15, rb = tid >> 4;
    uint32_t ra[8], rbw[TK / 16];
    auto load_chunk = [&](int n0) {
        const int j = n0 + 2 * p;
        BnCol c0, c1; int ch0 = 0, ch1 = 0;
        if (j < N) { ch0 = colmap_c(a.cm, j); c0 = load_bncol(a.tb, a.ldo, t, ch0, inv_n); }
        if (j + 1 < N) { ch1 = colmap_c(a.cm, j + 1); c1 = load_bncol(a.tb, a.ldo, t, ch1, inv_n); }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = row0 + rb + 16 * i;
            ra[i] = (r < a.Rt && j < N) ? load_dr_pair(a, (size_t)t * a.Rt + r, j, N, c0, c1, ch0, ch1) : 0u;
        }
#pragma unroll
        for (int i = 0; i < TK / 16; ++i) {
            const int kk = k0t + rb + 16 * i;
            rbw[i] = (rb + 16 * i < kblk * 8 && j < pa.Np) ? *reinterpret_cast<const uint32_t*>(pa.wn + (size_t)kk * pa.Np + j) : 0u;
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint32_t*>(As + (size_t)buf * kMmTM * kMmLd + (rb + 16 * i) * kMmLd + 2 * p) = ra[i];
#pragma unroll
        for (int i = 0; i < TK / 16; ++i)
            *reinterpret_cast<uint32_t*>(Bs + (size_t)buf * TK * kMmLd + (rb + 16 * i) * kMmLd + 2 * p) = rbw[i];
    };
    const int nchunks = (N + kMmKC - 1) / kMmKC;
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int nc = 0; nc < nchunks; ++nc) {
        const int buf = nc & 1;
        if (nc + 1 < nchunks) load_chunk((nc + 1) * kMmKC);
        const bf16* Ab = As + (size_t)buf * kMmTM * kMmLd;
        const bf16* Bb = Bs + (size_t)buf * TK * kMmLd;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t af[4];
            ldsm_x4(af, Ab + (warp * 16 + (lane & 15)) * kMmLd + ks * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int nb2 = 0; nb2 < NB / 2; ++nb2) {
                if (nb2 * 2 < kblk) {
                    uint32_t bf[4];
                    const int mi = lane >> 3;
                    ldsm_x4(bf, Bb + (nb2 * 16 + (mi >> 1) * 8 + (lane & 7)) * kMmLd + ks * 16 + (mi & 1) * 8);
                    mma_bf16(acc[nb2 * 2], af, bf[0], bf[1]);
                    mma_bf16(acc[nb2 * 2 + 1], af, bf[2], bf[3]);
                }
            }
        }
        if (nc + 1 < nchunks) store_chunk(buf ^ 1);
        __syncthreads();
    }
    bf16* Cs = smem;
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        if (nb < kblk) {
            const int j = nb * 8 + 2 * tg;
            *reinterpret_cast<uint32_t*>(Cs + (warp * 16 + g) * kCLd + j) = pack_bf16(acc[nb][0], acc[nb][1]);
            *reinterpret_cast<uint32_t*>(Cs + (warp * 16 + g + 8) * kCLd + j) = pack_bf16(acc[nb][2], acc[nb][3]);
        }
    }
    __syncthreads();
    const int rows = min(kMmTM, a.Rt - row0);
    {   // thread <-> fixed column pair, row lanes stride the rows (K is even for every layer of the tower)
        constexpr int kPairs = TK / 2, kRowLanes = 256 / kPairs;
        const int k = (tid % kPairs) * 2, rl = tid / kPairs;
        if (k < kcols) {
            bf16* dcol = a.dx + ((size_t)t * a.Rt + row0) * a.ldx + a.coffx + k0t + k;
            for (int r = rl; r < rows; r += kRowLanes) {
                bf16* d = dcol + (size_t)r * a.ldx;
                const uint32_t u = *reinterpret_cast<const uint32_t*>(Cs + r * kCLd + k);
                if (a.accumulate) {
                    float2 v = unpack_bf16(u);
                    const float2 o = unpack_bf16(*reinterpret_cast<const uint32_t*>(d));
                    *reinterpret_cast<uint32_t*>(d) = pack_bf16(v.x + o.x, v.y + o.y);
                } else *reinterpret_cast<uint32_t*>(d) = u;
            }
        }
    }
}

// ======================================================================================== weight gradient
// dW tile 64(k) x 64(j) per CTA; the 8 warps are 4 k-groups x 2 row halves of every 32-row chunk, the halves are
// folded through shared memory before the (fp32) atomics; rows are split over many CTAs (grid.z) for parallelism.
constexpr int kWgKT = 64, kWgNT = 64, kWgMC = 32, kWgLdX = kWgKT + 8, kWgLdR = kWgNT + 8;

CDRA_KERNEL __launch_bounds__(256, 3) pw_wgrad_mma_kernel(PwBwdArgs<bf16> a) {
    __shared__ __align__(16) bf16 Xs[2][kWgMC][kWgLdX];      // act(in) chunk [row][k]
    __shared__ __align__(16) bf16 Rs[2][kWgMC][kWgLdR];      // dR chunk      [row][j]
    __shared__ float red[4][32][33];                          // cross-half reduction
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kg = warp & 3, mh = warp >> 2;
    const int k0t = blockIdx.x * kWgKT, j0t = blockIdx.y * kWgNT;
    const int t = blockIdx.z / a.row_splits, sp = blockIdx.z % a.row_splits;
    const int N = a.cm.n, K = a.K;
    const double inv_n = 1.0 / (double)a.Rt;
    const int rows_per = ((a.Rt + a.row_splits - 1) / a.row_splits + kWgMC - 1) / kWgMC * kWgMC;
    const int rbeg = sp * rows_per, rend = min(a.Rt, rbeg + rows_per);
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    // loaders: pair p (2 columns), rows rq + 8 i
    const int p = tid & 31, rq = tid >> 5;
    const int kx = k0t + 2 * p, jr = j0t + 2 * p;
    float sx0 = 1.f, hx0 = 0.f, sx1 = 1.f, hx1 = 0.f;
    if (a.in.aff) {
        if (kx < K) { const float2 f = a.in.aff[(size_t)t * a.in.ld + a.in.coff + kx]; sx0 = f.x; hx0 = f.y; }
        if (kx + 1 < K) { const float2 f = a.in.aff[(size_t)t * a.in.ld + a.in.coff + kx + 1]; sx1 = f.x; hx1 = f.y; }
    }
    BnCol c0, c1; int ch0 = 0, ch1 = 0;
    if (jr < N) { ch0 = colmap_c(a.cm, jr); c0 = load_bncol(a.tb, a.ldo, t, ch0, inv_n); }
    if (jr + 1 < N) { ch1 = colmap_c(a.cm, jr + 1); c1 = load_bncol(a.tb, a.ldo, t, ch1, inv_n); }
    const bf16* in = (const bf16*)a.in.data;
    uint32_t vx[4], vr[4];
    auto load_chunk = [&](int m0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = m0 + rq + 8 * i;
            uint32_t v = 0u;
            if (r < rend) {
                if (kx < K) {
                    float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(in + ((size_t)t * a.Rt + r) * a.in.ld + a.in.coff + kx));
                    f.x = fmaf(f.x, sx0, hx0); f.y = fmaf(f.y, sx1, hx1);
                    if (a.in.clamp) { f.x = relu6f(f.x); f.y = relu6f(f.y); }
                    v = pack_bf16(f.x, f.y);
                } else if (kx == K) v = pack_bf16(1.f, 0.f);          // virtual ones row -> bias gradient
            }
            vx[i] = v;
            vr[i] = (r < rend && jr < N) ? load_dr_pair(a, (size_t)t * a.Rt + r, jr, N, c0, c1, ch0, ch1) : 0u;
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<uint32_t*>(&Xs[buf][rq + 8 * i][2 * p]) = vx[i];
            *reinterpret_cast<uint32_t*>(&Rs[buf][rq + 8 * i][2 * p]) = vr[i];
        }
    };
    const int nchunks = rend > rbeg ? (rend - rbeg + kWgMC - 1) / kWgMC : 0;
    if (nchunks > 0) { load_chunk(rbeg); store_chunk(0); }
    __syncthreads();
    for (int mc = 0; mc < nchunks; ++mc) {
        const int buf = mc & 1;
        if (mc + 1 < nchunks) load_chunk(rbeg + (mc + 1) * kWgMC);
        {
            uint32_t af[4];
            const int mi = lane >> 3;
            ldsm_x4_trans(af, &Xs[buf][mh * 16 + (mi >> 1) * 8 + (lane & 7)][kg * 16 + (mi & 1) * 8]);
#pragma unroll
            for (int nb2 = 0; nb2 < 4; ++nb2) {
                uint32_t bf[4];
                ldsm_x4_trans(bf, &Rs[buf][mh * 16 + (mi & 1) * 8 + (lane & 7)][nb2 * 16 + (mi >> 1) * 8]);
                mma_bf16(acc[nb2 * 2], af, bf[0], bf[1]);
                mma_bf16(acc[nb2 * 2 + 1], af, bf[2], bf[3]);
            }
        }
        if (mc + 1 < nchunks) store_chunk(buf ^ 1);
        __syncthreads();
    }
    // fold the two row halves, then one atomic per output element
    if (mh == 1) {
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
#pragma unroll
            for (int e = 0; e < 4; ++e) red[kg][nb * 4 + e][lane] = acc[nb][e];
    }
    __syncthreads();
    // partial 64x64 tile of this CTA -> scratch (plain stores); pw_wgrad_reduce_kernel sums the row splits.
    // (atomics from ~1000 CTAs onto the same 4096 addresses cost ~150 us per launch, and are not deterministic)
    if (mh == 0) {
        float* part = a.partials + (((size_t)blockIdx.z * gridDim.x + blockIdx.x) * gridDim.y + blockIdx.y) * (kWgKT * kWgNT);
        const int g = lane >> 2, tg = lane & 3;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int jl = nb * 8 + 2 * tg;
            const float2 lo = make_float2(acc[nb][0] + red[kg][nb * 4 + 0][lane], acc[nb][1] + red[kg][nb * 4 + 1][lane]);
            const float2 hi = make_float2(acc[nb][2] + red[kg][nb * 4 + 2][lane], acc[nb][3] + red[kg][nb * 4 + 3][lane]);
            *reinterpret_cast<float2*>(part + (kg * 16 + g) * kWgNT + jl) = lo;
            *reinterpret_cast<float2*>(part + (kg * 16 + g + 8) * kWgNT + jl) = hi;
        }
    }
}

// sums the per-CTA partial tiles (fixed order -> deterministic), writes dW / db / dgamma / dbeta
struct PwWgReduceArgs { PwBwdArgs<bf16> a; int kt, nt, nz; };
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) pw_wgrad_reduce_kernel(PwWgReduceArgs ra) {
    const PwBwdArgs<bf16>& a = ra.a;
    // grid = (kt, nt, 128): each CTA owns 32 elements of one tile; its 8 warps split the row-split partials
    __shared__ float red[8][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, N = a.cm.n, K = a.K;
    const int k0t = blockIdx.x * kWgKT, j0t = blockIdx.y * kWgNT;
    const size_t tile = (size_t)kWgKT * kWgNT, zstride = (size_t)ra.kt * ra.nt * tile;
    const float* base = a.partials + ((size_t)blockIdx.x * ra.nt + blockIdx.y) * tile;
    const int e = blockIdx.z * 32 + lane;
    float s = 0.f;
    for (int z = warp; z < ra.nz; z += 8) s += base[(size_t)z * zstride + e];
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
        s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][lane];
        const int k = k0t + e / kWgNT, j = j0t + e % kWgNT;
        if (k <= K && j < N) {
            const int wc = colmap_w(a.cm, j);
            if (k < K) a.dw[(size_t)k * N + wc] = s; else a.db[wc] = s;
        }
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {       // BN parameter gradients
        for (int j = tid; j < N; j += 256) {
            const int c = colmap_c(a.cm, j), wc = colmap_w(a.cm, j);
            double gs = 0.0, bs = 0.0;
            for (int tt = 0; tt < kT; ++tt) { const double2 s = a.tb.bst[(size_t)tt * a.ldo + c]; bs += s.x; gs += s.y; }
            a.dgamma[wc] = (float)gs; a.dbeta[wc] = (float)bs;
        }
    }
}

}  // namespace cdra
#endif  // CDRA_EMU
