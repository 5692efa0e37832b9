// fp32 kernels for everything after the image tower: feature MLPs (core/architectures.py:9-27),
// GRUs + trunk (core/networks.py:24-56), control branches (core/networks.py:59-66), generic dense
// GEMMs used by their forward/backward.
#pragma once
#include "cdra_common.cuh"

namespace cdra {

// --------------------------------------------------------------------------- generic fp32 GEMM
//   C[M][N] (=|+=) opA(A) * opB(B) (+ bias[n]);  TA: A stored [K][M];  TB: B stored [N][K]
struct GemmArgs {
    const float* A; int lda;
    const float* B; int ldb;
    float* C; int ldc;
    const float* bias;
    int M, N, K, accumulate;
};
constexpr int kGT = 64, kGK = 16;

template <bool TA, bool TB>
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) sgemm_kernel(GemmArgs a) {
    CDRA_SHARED float As[kGK][kGT + 4];
    CDRA_SHARED float Bs[kGK][kGT + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * kGT, n0 = blockIdx.y * kGT;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < a.K; k0 += kGK) {
        if (!TA) {
            const int m = m0 + (tid >> 2), kq = (tid & 3) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + kq + q;
                As[kq + q][tid >> 2] = (m < a.M && k < a.K) ? a.A[(size_t)m * a.lda + k] : 0.f;
            }
        } else {
            const int k = k0 + (tid >> 4), mq = (tid & 15) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m = m0 + mq + q;
                As[tid >> 4][mq + q] = (m < a.M && k < a.K) ? a.A[(size_t)k * a.lda + m] : 0.f;
            }
        }
        if (!TB) {
            const int k = k0 + (tid >> 4), nq = (tid & 15) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int n = n0 + nq + q;
                Bs[tid >> 4][nq + q] = (n < a.N && k < a.K) ? a.B[(size_t)k * a.ldb + n] : 0.f;
            }
        } else {
            const int n = n0 + (tid >> 2), kq = (tid & 3) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + kq + q;
                Bs[kq + q][tid >> 2] = (n < a.N && k < a.K) ? a.B[(size_t)n * a.ldb + k] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kGK; ++kk) {
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
        const int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
            float* c = a.C + (size_t)m * a.ldc + n;
            *c = a.accumulate ? *c + v : v;
        }
    }
}

#ifndef CDRA_EMU
// --------------------------------------------------------------------------- the same GEMM on tensor cores (perf mode)
// TF32 mma.sync m16n8k8, fp32 accumulate: operands are fp32 in HBM and are rounded to TF32 (10-bit mantissa) when they
// are staged; used by the bf16 "perf mode" for the GRU projections, the trunk and the control branches (the fp32 parity
// mode keeps sgemm_kernel).  64x64x32 tiles, register-prefetched double buffer; each operand keeps its HBM orientation
// in shared memory (coalesced, conflict-free staging) and the fragment indexing follows the orientation.
constexpr int kTgK = 32;
CDRA_DEV uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
CDRA_DEV void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// stage a [rows x cols] block of a row-major matrix (row stride ld) starting at (r0, c0) into registers: 8 values / thread
template <int ROWS, int COLS>
CDRA_DEV void tg_load(float (&v)[8], const float* M, int ld, int r0, int c0, int nrows, int ncols, bool vec) {
    // ROWS x COLS = 2048 values, 256 threads, 2 float4 per thread: thread covers rows (tid / (COLS/4)) + {0, ROWS/2}
    const int cq = (threadIdx.x % (COLS / 4)) * 4, rr = threadIdx.x / (COLS / 4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = r0 + rr + h * (ROWS / 2), c = c0 + cq;
        if (r < nrows && c + 3 < ncols && vec) {
            const float4 t = *reinterpret_cast<const float4*>(M + (size_t)r * ld + c);
            v[4 * h] = t.x; v[4 * h + 1] = t.y; v[4 * h + 2] = t.z; v[4 * h + 3] = t.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[4 * h + q] = (r < nrows && c + q < ncols) ? M[(size_t)r * ld + c + q] : 0.f;
        }
    }
}
template <int ROWS, int COLS, int LD>
CDRA_DEV void tg_store(uint32_t* S, const float (&v)[8]) {
    const int cq = (threadIdx.x % (COLS / 4)) * 4, rr = threadIdx.x / (COLS / 4);
#pragma unroll
    for (int h = 0; h < 2; ++h)
        *reinterpret_cast<uint4*>(S + (rr + h * (ROWS / 2)) * LD + cq) =
            make_uint4(to_tf32(v[4 * h]), to_tf32(v[4 * h + 1]), to_tf32(v[4 * h + 2]), to_tf32(v[4 * h + 3]));
}

struct TGemmArgs { GemmArgs g; int vecA, vecB; };

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) tgemm_kernel(const TGemmArgs ta) {
    const GemmArgs& a = ta.g;
    // A: !TA -> [64 m][32 k + 4] ; TA -> [32 k][64 m + 8].   B: !TB -> [32 k][64 n + 8] ; TB -> [64 n][32 k + 4]   (2304 words each)
    __shared__ __align__(16) uint32_t As[2][2304];
    __shared__ __align__(16) uint32_t Bs[2][2304];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;             // warp tile: rows wm*16.., columns wn*32..
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    float ra[8], rb[8];
    auto load = [&](int k0) {
        if (!TA) tg_load<64, 32>(ra, a.A, a.lda, m0, k0, a.M, a.K, ta.vecA != 0);
        else tg_load<32, 64>(ra, a.A, a.lda, k0, m0, a.K, a.M, ta.vecA != 0);
        if (!TB) tg_load<32, 64>(rb, a.B, a.ldb, k0, n0, a.K, a.N, ta.vecB != 0);
        else tg_load<64, 32>(rb, a.B, a.ldb, n0, k0, a.N, a.K, ta.vecB != 0);
    };
    auto store = [&](int buf) {
        if (!TA) tg_store<64, 32, 36>(As[buf], ra); else tg_store<32, 64, 72>(As[buf], ra);
        if (!TB) tg_store<32, 64, 72>(Bs[buf], rb); else tg_store<64, 32, 36>(Bs[buf], rb);
    };
    const int nk = (a.K + kTgK - 1) / kTgK;
    load(0); store(0);
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load((kb + 1) * kTgK);
        const uint32_t* A_ = As[buf]; const uint32_t* B_ = Bs[buf];
#pragma unroll
        for (int ks = 0; ks < kTgK; ks += 8) {
            uint32_t af[4];
            const int r = wm * 16 + g;
            if (!TA) { af[0] = A_[r * 36 + ks + tg]; af[1] = A_[(r + 8) * 36 + ks + tg]; af[2] = A_[r * 36 + ks + tg + 4]; af[3] = A_[(r + 8) * 36 + ks + tg + 4]; }
            else { af[0] = A_[(ks + tg) * 72 + r]; af[1] = A_[(ks + tg) * 72 + r + 8]; af[2] = A_[(ks + tg + 4) * 72 + r]; af[3] = A_[(ks + tg + 4) * 72 + r + 8]; }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const int n = wn * 32 + nb * 8 + g;
                uint32_t b0, b1;
                if (!TB) { b0 = B_[(ks + tg) * 72 + n]; b1 = B_[(ks + tg + 4) * 72 + n]; }
                else { b0 = B_[n * 36 + ks + tg]; b1 = B_[n * 36 + ks + tg + 4]; }
                mma_tf32(acc[nb], af, b0, b1);
            }
        }
        if (kb + 1 < nk) store(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int m = m0 + wm * 16 + g + (e >> 1) * 8, n = n0 + wn * 32 + nb * 8 + 2 * tg + (e & 1);
            if (m < a.M && n < a.N) {
                const float v = acc[nb][e] + (a.bias ? a.bias[n] : 0.f);
                float* c = a.C + (size_t)m * a.ldc + n;
                *c = a.accumulate ? *c + v : v;
            }
        }
}
#endif

// out[n] (=|+=) sum_m X[m][n]
struct ColsumArgs { const float* X; int ldx, M, N; float* out; int accumulate; };
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) colsum_kernel(ColsumArgs a) {
    CDRA_SHARED double red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + cx;
    double s = 0.0;
    if (n < a.N) for (int m = ry; m < a.M; m += 8) s += (double)a.X[(size_t)m * a.ldx + n];
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && n < a.N) {
        for (int i = 1; i < 8; ++i) s += red[i][cx];
        a.out[n] = a.accumulate ? a.out[n] + (float)s : (float)s;
    }
}

// --------------------------------------------------------------------------- BatchNorm over a [B][C] matrix
struct Bn1dArgs {
    const float* x; int ldx;
    float* y; int ldy;            // forward output / backward dx
    const float* dy; int lddy;    // backward only
    float2* stat;                 // [C] (mean, inv_std)
    const float* gamma; const float* beta;
    float* mov_mean; float* mov_var;
    float* dgamma; float* dbeta;
    int B, C, training;
};

CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) bn1d_fwd_kernel(Bn1dArgs a) {
    CDRA_SHARED double r1[8][33], r2[8][33];
    CDRA_SHARED float s_mean[32], s_inv[32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    if (a.training) {
        double s = 0.0, q = 0.0;
        if (c < a.C) {
#pragma unroll 8
            for (int b = ry; b < a.B; b += 8) { const double v = a.x[(size_t)b * a.ldx + c]; s += v; q += v * v; }
        }
        r1[ry][cx] = s; r2[ry][cx] = q;
        __syncthreads();
        if (ry == 0 && c < a.C) {
            for (int i = 1; i < 8; ++i) { s += r1[i][cx]; q += r2[i][cx]; }
            const double mean = s / a.B;
            double var = q / a.B - mean * mean;
            if (var < 0.0) var = 0.0;
            const float inv = (float)(1.0 / sqrt(var + (double)kBnEps));
            s_mean[cx] = (float)mean; s_inv[cx] = inv;
            a.stat[c] = make_float2((float)mean, inv);
            if (a.mov_mean) {   // 2-D input: Keras non-fused path feeds the biased variance
                a.mov_mean[c] -= (a.mov_mean[c] - (float)mean) * (1.f - kBnMomentum);
                a.mov_var[c] -= (a.mov_var[c] - (float)var) * (1.f - kBnMomentum);
            }
        }
    } else if (ry == 0 && c < a.C) {
        const float inv = (float)(1.0 / sqrt((double)a.mov_var[c] + (double)kBnEps));
        s_mean[cx] = a.mov_mean[c]; s_inv[cx] = inv;
        a.stat[c] = make_float2(a.mov_mean[c], inv);
    }
    __syncthreads();
    if (c < a.C) {
        const float g = a.gamma[c] * s_inv[cx], sh = a.beta[c] - s_mean[cx] * g;
        // eight rows per round, loads first: a store that depends on a load, followed by loads that may alias it, costs a memory
        // round trip per row otherwise (64 rows per thread)
        const float* CDRA_RESTRICT xp = a.x + c; float* CDRA_RESTRICT yp = a.y + c;
        for (int b0 = ry; b0 < a.B; b0 += 64) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int b = b0 + 8 * u; v[u] = b < a.B ? xp[(size_t)b * a.ldx] : 0.f; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int b = b0 + 8 * u; if (b < a.B) yp[(size_t)b * a.ldy] = fmaf(v[u], g, sh); }
        }
    }
}

CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) bn1d_bwd_kernel(Bn1dArgs a) {
    CDRA_SHARED double r1[8][33], r2[8][33];
    CDRA_SHARED float s_k1[32], s_k2[32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float mean = 0.f, inv = 0.f;
    if (c < a.C) { const float2 st = a.stat[c]; mean = st.x; inv = st.y; }
    double s = 0.0, q = 0.0;
    if (c < a.C) {
#pragma unroll 8
        for (int b = ry; b < a.B; b += 8) {
            const double d = a.dy[(size_t)b * a.lddy + c];
            s += d; q += d * (double)((a.x[(size_t)b * a.ldx + c] - mean) * inv);
        }
    }
    r1[ry][cx] = s; r2[ry][cx] = q;
    __syncthreads();
    if (ry == 0 && c < a.C) {
        for (int i = 1; i < 8; ++i) { s += r1[i][cx]; q += r2[i][cx]; }
        a.dgamma[c] = (float)q; a.dbeta[c] = (float)s;
        s_k1[cx] = (float)(s / a.B); s_k2[cx] = (float)(q / a.B);
    }
    __syncthreads();
    if (c < a.C) {
        const float g = a.gamma[c] * inv, k1 = s_k1[cx], k2 = s_k2[cx];
        const float* CDRA_RESTRICT xp = a.x + c; const float* CDRA_RESTRICT dp = a.dy + c; float* CDRA_RESTRICT yp = a.y + c;
        for (int b0 = ry; b0 < a.B; b0 += 64) {              // eight rows per round, loads first (see bn1d_fwd_kernel)
            float xv[8], dv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int b = b0 + 8 * u;
                xv[u] = b < a.B ? xp[(size_t)b * a.ldx] : 0.f; dv[u] = b < a.B ? dp[(size_t)b * a.lddy] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int b = b0 + 8 * u;
                if (b < a.B) { const float xh = (xv[u] - mean) * inv; yp[(size_t)b * a.ldy] = g * (dv[u] - k1 - xh * k2); }
            }
        }
    }
}

// --------------------------------------------------------------------------- swish6 (rl/utils.py:420-421)
CDRA_DEV float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
struct ActArgs { const float* x; const float* dy; float* y; long long n; };
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) swish6_fwd_kernel(ActArgs a) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n) return;
    const float x = a.x[i];
    a.y[i] = fminf(x * sigmoidf_(x), 6.f);
}
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) swish6_bwd_kernel(ActArgs a) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n) return;
    const float x = a.x[i], s = sigmoidf_(x);
    a.y[i] = (x * s <= 6.f) ? a.dy[i] * (s * (1.f + x * (1.f - s))) : 0.f;
}

// --------------------------------------------------------------------------- GRU gates (Keras reset_after, [z|r|h])
struct GruGateArgs {
    const float* xp;        // [B][3u] = x W + b0
    float* hp;              // [B][3u] = h_prev R + b1   (written from b1 when hprev == nullptr)
    const float* b1;        // recurrent bias row (used when hprev == nullptr)
    const float* hprev; int ldhp;    // [B][u] or nullptr (h0 = 0)
    float* hout; int ldho;           // [B][u]
    // backward
    const float* dh; int lddh;       // d loss / d h_t
    float* dxp; float* dhp;          // [B][3u]
    float* dhprev; int lddhp;        // [B][u]  (= dh * z; the recurrent GEMM accumulates on top)
    int B, u;
};
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) gru_gate_fwd_kernel(GruGateArgs a) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)a.B * a.u) return;
    const int j = (int)(i % a.u), b = (int)(i / a.u);
    const size_t o = (size_t)b * 3 * a.u;
    float hz, hr, hh, hpv = 0.f;
    if (a.hprev) { hz = a.hp[o + j]; hr = a.hp[o + a.u + j]; hh = a.hp[o + 2 * a.u + j]; hpv = a.hprev[(size_t)b * a.ldhp + j]; }
    else {
        hz = a.b1[j]; hr = a.b1[a.u + j]; hh = a.b1[2 * a.u + j];
        a.hp[o + j] = hz; a.hp[o + a.u + j] = hr; a.hp[o + 2 * a.u + j] = hh;
    }
    const float z = sigmoidf_(a.xp[o + j] + hz);
    const float r = sigmoidf_(a.xp[o + a.u + j] + hr);
    const float hc = tanhf(a.xp[o + 2 * a.u + j] + r * hh);
    a.hout[(size_t)b * a.ldho + j] = z * hpv + (1.f - z) * hc;
}
CDRA_KERNEL CDRA_LAUNCH_BOUNDS(256) gru_gate_bwd_kernel(GruGateArgs a) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)a.B * a.u) return;
    const int j = (int)(i % a.u), b = (int)(i / a.u);
    const size_t o = (size_t)b * 3 * a.u;
    const float hz = a.hp[o + j], hr = a.hp[o + a.u + j], hh = a.hp[o + 2 * a.u + j];
    const float hpv = a.hprev ? a.hprev[(size_t)b * a.ldhp + j] : 0.f;
    const float z = sigmoidf_(a.xp[o + j] + hz);
    const float r = sigmoidf_(a.xp[o + a.u + j] + hr);
    const float hc = tanhf(a.xp[o + 2 * a.u + j] + r * hh);
    const float dh = a.dh[(size_t)b * a.lddh + j];
    const float dz = dh * (hpv - hc), dhc = dh * (1.f - z);
    const float dah = dhc * (1.f - hc * hc);
    const float dr = dah * hh;
    const float daz = dz * z * (1.f - z), dar = dr * r * (1.f - r);
    a.dxp[o + j] = daz; a.dxp[o + a.u + j] = dar; a.dxp[o + 2 * a.u + j] = dah;
    a.dhp[o + j] = daz; a.dhp[o + a.u + j] = dar; a.dhp[o + 2 * a.u + j] = dah * r;
    if (a.dhprev) a.dhprev[(size_t)b * a.lddhp + j] = dh * z;
}

// --------------------------------------------------------------------------- feature MLPs (one CTA per modality)
// per slice: Dense(d->16, relu6) -> BN -> Dense(16->16, relu6) -> BN  (core/architectures.py:9-27,
// called with units=16, num_layers=2, activation=relu6, core/carla_agent.py:62-64)
struct FeatArgs {
    const float* x; int d;                  // [B][kT][d]
    const float *w1, *b1, *g1, *be1, *w2, *b2, *g2, *be2;
    float *mm1, *mv1, *mm2, *mv2;
    float *h1, *h2, *n1, *out;              // [kT][B][16]
    float2 *st1, *st2;                      // [kT][16] (mean, inv)
    // backward
    float* dout;                            // [kT][B][16] in: d loss/d out ; reused as scratch
    float* dtmp;                            // [kT][B][16] scratch
    float *dw1, *db1, *dg1, *dbe1, *dw2, *db2, *dg2, *dbe2;
};
struct FeatArgs3 { FeatArgs m[3]; int B, training; };
constexpr int kFU = 16, kFeatThreads = 512, kFeatLanes = kFeatThreads / kFU;

CDRA_DEV double feat_reduce(double v, int n, int lane, double (*red)[kFU]) {
    __syncthreads();
    red[lane][n] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < kFeatLanes; ++i) s += red[i][n];
    return s;       // every thread gets the total of its column n
}

CDRA_KERNEL CDRA_LAUNCH_BOUNDS(kFeatThreads) featnet_fwd_kernel(FeatArgs3 aa) {
    CDRA_SHARED double red[kFeatLanes][kFU];
    CDRA_SHARED float sw1[kFU * kFU], sw2[kFU * kFU];
    const FeatArgs& a = aa.m[blockIdx.x];
    const int B = aa.B, tid = threadIdx.x, n = tid % kFU, lane = tid / kFU, d = a.d;
    for (int i = tid; i < d * kFU; i += kFeatThreads) sw1[i] = a.w1[i];
    for (int i = tid; i < kFU * kFU; i += kFeatThreads) sw2[i] = a.w2[i];
    __syncthreads();
    float mm1 = a.mm1 ? a.mm1[n] : 0.f, mv1 = a.mv1 ? a.mv1[n] : 1.f, mm2 = a.mm2 ? a.mm2[n] : 0.f, mv2 = a.mv2 ? a.mv2[n] : 1.f;
    for (int t = 0; t < kT; ++t) {
        for (int layer = 0; layer < 2; ++layer) {
            const float* W = layer ? sw2 : sw1;
            const float bias = layer ? a.b2[n] : a.b1[n];
            const int K = layer ? kFU : d;
            float* hbuf = (layer ? a.h2 : a.h1) + (size_t)t * B * kFU;
            double s = 0.0, q = 0.0;
            for (int b = lane; b < B; b += kFeatLanes) {
                const float* xin = layer ? a.n1 + ((size_t)t * B + b) * kFU : a.x + ((size_t)b * kT + t) * d;
                float v = bias;
                for (int k = 0; k < K; ++k) v = fmaf(xin[k], W[k * kFU + n], v);
                v = relu6f(v);
                hbuf[(size_t)b * kFU + n] = v;
                s += v; q += (double)v * v;
            }
            float mean, inv;
            float& mm = layer ? mm2 : mm1; float& mv = layer ? mv2 : mv1;
            if (aa.training) {
                s = feat_reduce(s, n, lane, red);
                q = feat_reduce(q, n, lane, red);
                const double m_ = s / B; double var = q / B - m_ * m_; if (var < 0.0) var = 0.0;
                mean = (float)m_; inv = (float)(1.0 / sqrt(var + (double)kBnEps));
                mm -= (mm - mean) * (1.f - kBnMomentum);
                mv -= (mv - (float)var) * (1.f - kBnMomentum);
            } else { mean = mm; inv = (float)(1.0 / sqrt((double)mv + (double)kBnEps)); }
            if (lane == 0) (layer ? a.st2 : a.st1)[t * kFU + n] = make_float2(mean, inv);
            const float g = (layer ? a.g2[n] : a.g1[n]) * inv, sh = (layer ? a.be2[n] : a.be1[n]) - mean * g;
            float* obuf = (layer ? a.out : a.n1) + (size_t)t * B * kFU;
            for (int b = lane; b < B; b += kFeatLanes) obuf[(size_t)b * kFU + n] = fmaf(hbuf[(size_t)b * kFU + n], g, sh);
            __syncthreads();
        }
    }
    if (aa.training && lane == 0 && a.mm1) { a.mm1[n] = mm1; a.mv1[n] = mv1; a.mm2[n] = mm2; a.mv2[n] = mv2; }
}

CDRA_KERNEL CDRA_LAUNCH_BOUNDS(kFeatThreads) featnet_bwd_kernel(FeatArgs3 aa) {
    CDRA_SHARED double red[kFeatLanes][kFU];
    CDRA_SHARED float sw2[kFU * kFU];
    const FeatArgs& a = aa.m[blockIdx.x];
    const int B = aa.B, tid = threadIdx.x, n = tid % kFU, lane = tid / kFU, d = a.d;
    for (int i = tid; i < kFU * kFU; i += kFeatThreads) sw2[i] = a.w2[i];
    __syncthreads();
    double dg1 = 0, dbe1 = 0, dg2 = 0, dbe2 = 0, db1 = 0, db2 = 0;
    float dw1 = 0.f, dw2 = 0.f;            // thread tid < K*16 owns dW[k = tid/16][n]
    for (int t = 0; t < kT; ++t) {
        const size_t so = (size_t)t * B * kFU;
        for (int layer = 1; layer >= 0; --layer) {
            const float2 st = (layer ? a.st2 : a.st1)[t * kFU + n];
            const float* hbuf = (layer ? a.h2 : a.h1) + so;
            float* din = a.dout + so;        // gradient wrt this BN's output ([b][n])
            double s1 = 0.0, s2 = 0.0;
            for (int b = lane; b < B; b += kFeatLanes) {
                const double dv = din[(size_t)b * kFU + n];
                s1 += dv; s2 += dv * (double)((hbuf[(size_t)b * kFU + n] - st.x) * st.y);
            }
            s1 = feat_reduce(s1, n, lane, red);
            s2 = feat_reduce(s2, n, lane, red);
            if (layer) { dg2 += s2; dbe2 += s1; } else { dg1 += s2; dbe1 += s1; }
            const float g = (layer ? a.g2[n] : a.g1[n]) * st.y, k1 = (float)(s1 / B), k2 = (float)(s2 / B);
            float* dp = a.dtmp + so;         // gradient wrt the dense pre-activation
            double sb = 0.0;
            for (int b = lane; b < B; b += kFeatLanes) {
                const float h = hbuf[(size_t)b * kFU + n];
                const float dh = g * (din[(size_t)b * kFU + n] - k1 - (h - st.x) * st.y * k2);
                const float v = (h > 0.f && h < 6.f) ? dh : 0.f;
                dp[(size_t)b * kFU + n] = v;
                sb += v;
            }
            sb = feat_reduce(sb, n, lane, red);
            if (layer) db2 += sb; else db1 += sb;
            __syncthreads();
            const int K = layer ? kFU : d;
            if (tid < K * kFU) {             // weight gradient: dW[k][n] += sum_b in[b][k] * dp[b][n]
                const int k = tid / kFU;
                float acc = 0.f;
                for (int b = 0; b < B; ++b) {
                    const float xin = layer ? a.n1[so + (size_t)b * kFU + k] : a.x[((size_t)b * kT + t) * d + k];
                    acc = fmaf(xin, dp[(size_t)b * kFU + n], acc);
                }
                if (layer) dw2 += acc; else dw1 += acc;
            }
            if (layer) {                     // d n1[b][k] = sum_n dp[b][n] * W2[k][n]   (k == this thread's n)
                for (int b = lane; b < B; b += kFeatLanes) {
                    float acc = 0.f;
                    for (int j = 0; j < kFU; ++j) acc = fmaf(dp[(size_t)b * kFU + j], sw2[n * kFU + j], acc);
                    din[(size_t)b * kFU + n] = acc;
                }
            }
            __syncthreads();
        }
    }
    if (tid < kFU * kFU) a.dw2[tid] = dw2;
    if (tid < d * kFU) a.dw1[tid] = dw1;
    if (lane == 0) {
        a.dg1[n] = (float)dg1; a.dbe1[n] = (float)dbe1; a.dg2[n] = (float)dg2; a.dbe2[n] = (float)dbe2;
        a.db1[n] = (float)db1; a.db2[n] = (float)db2;
    }
}

}  // namespace cdra
