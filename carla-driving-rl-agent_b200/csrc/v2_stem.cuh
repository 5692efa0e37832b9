// v2 tower: the stem (core/architectures.py:153-161) for uint8 frames --
//     Conv 3x3 s2 valid 3->24 (+bias) on u8/255, BatchNorm, ReLU6, MaxPool 3x3 s2 SAME
// as three band kernels (a band = a run of output rows of one frame, sized so 2+ CTAs fit per SM):
//   stem_fwd_kernel  u8 frame rows by TMA -> im2col fragments gathered straight from shared memory ->
//                    mma.sync m16n8k16 (A = integer pixel values, exact in bf16; B = W/255) -> bf16 raw rows staged in
//                    shared memory -> one TMA bulk store; per-(slice, channel) sums of the stored values
//   pool_fwd_kernel  raw stem rows by TMA, BN affine + ReLU6 in place, 3x3 window max (first maximum in row-major
//                    order wins, like TF's MaxPoolGrad scan) -> pool output + the winner's window position (u8)
//   stem_bwd_kernel  the whole stem backward in ONE pass over the data: max-pool backward (scatter through the
//                    stored positions), ReLU6 mask, and the weight gradient with the BatchNorm backward folded in by
//                    linearity:   dW[k][c] = scale_c/255 * ( G1[k][c] - S1_c/n * G0[k] - S2_c/n * G2[k][c] )
//                    G1 = sum P'[r][k] dz[r][c],  G2 = sum P'[r][k] xhat[r][c],  G0 = sum P'[r][k],  P' = pixel - 128
//                    (centred so the three terms are fluctuation-sized, no cancellation), S1 = sum dz, S2 = sum dz xhat.
//                    G1 / G2 / G0 are tensor-core products with fragments built in registers.
//   stem_bwd_finish  combines the per-slice sums into dW, dgamma, dbeta (the conv bias feeds a training-mode
//                    BatchNorm: its gradient is analytically zero).
#pragma once
#ifndef CDRA_EMU
#include "v2_common.cuh"

namespace cdra {
namespace v2 {

constexpr int kStemThreads = 256, kStemWarps = 8, kSC = 24;       // kSC == kStemC
constexpr int kGaccN = 32 * kSC * 2 + 32 + 2 * kSC;               // G1[32][24], G2[32][24], G0[32], S1[24], S2[24]

struct StemGeom { int B, H, W, W3, Hs, Ws, Hp, Wp, pad_t, pad_l; };

CDRA_DEV void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
CDRA_DEV void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
CDRA_DEV void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
CDRA_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
CDRA_DEV void red_bf16x2(void* smem_ptr, uint32_t v) {
    asm volatile("red.shared.add.noftz.bf16x2 [%0], %1;" ::"r"(smem_u32(smem_ptr)), "r"(v) : "memory");
}

// pixel table of a band: entry q (row-major inside the band) = image byte offset of the patch origin relative to
// the band's first image row (low 16 bits) | x << 16 | relative row << 24
CDRA_DEV void stem_fill_table(uint32_t* tbl, int npx_max, const StemGeom& g) {
    for (int q = threadIdx.x; q < npx_max; q += blockDim.x) {
        const int ry = q / g.Ws, x = q - ry * g.Ws;
        tbl[q] = (uint32_t)(ry * 2 * g.W3 + x * 6) | ((uint32_t)x << 16) | ((uint32_t)ry << 24);
    }
}
CDRA_DEV int stem_koff(int k, int W3) { return k < 27 ? (k / 9) * W3 + (k % 9) : 0; }

// ======================================================================================== stem forward
struct StemFwdArgs {
    const uint8_t* img;             // [B][kT][H][W][3]
    StemGeom g;
    const float* w; const float* bias;
    bf16* out;                      // [kT*B][Hs*Ws][24] raw conv output
    double2* fst;                   // replicated forward sums (BnTables.fst of the legacy stem tensor)
    int HB, nbands, units_per_cta, training;
};
struct StemFwdSmem { int tbl, stat, img, tile, total, img_stride; };
inline __host__ __device__ StemFwdSmem stem_fwd_smem(int HB, int Ws, int W3) {
    StemFwdSmem s; int off = 64;
    s.tbl = off; off += ((HB * Ws + 15) & ~15) * 4;
    s.stat = off; off += 2 * kSC * 4;
    off = (off + 127) & ~127;
    s.img_stride = (((2 * HB + 1) * W3 + 15 + 127) & ~127);
    s.img = off; off += 2 * s.img_stride;
    s.tile = off; off += HB * Ws * kSC * 2;
    s.total = (off + 127) & ~127;
    return s;
}

__global__ void __launch_bounds__(kStemThreads, 2) stem_fwd_kernel(const StemFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const StemGeom& G = a.g;
    const StemFwdSmem L = stem_fwd_smem(a.HB, G.Ws, G.W3);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint32_t* tbl = reinterpret_cast<uint32_t*>(smem + L.tbl);
    float* s_stat = reinterpret_cast<float*>(smem + L.stat);
    bf16* tile = reinterpret_cast<bf16*>(smem + L.tile);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const int nunits = kT * G.B * a.nbands;
    const int u_lo = blockIdx.x * a.units_per_cta, u_hi = min(nunits, u_lo + a.units_per_cta);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    stem_fill_table(tbl, a.HB * G.Ws, G);
    if (tid < 2 * kSC) s_stat[tid] = 0.f;
    __syncthreads();

    const size_t frame_bytes = (size_t)G.H * G.W3;
    auto issue = [&](int u, int buf) {                 // thread 0
        const int f = u / a.nbands, band = u - f * a.nbands, t = f / G.B, b = f - t * G.B;
        const int y0 = band * a.HB, n = min(a.HB, G.Hs - y0);
        const uint32_t bytes = (uint32_t)(((2 * n + 1) * G.W3 + 15) & ~15);
        mbar_expect_tx(&full[buf], bytes);
        bulk_g2s(smem + L.img + (size_t)buf * L.img_stride, a.img + (size_t)(b * kT + t) * frame_bytes + (size_t)2 * y0 * G.W3, bytes, &full[buf]);
    };
    if (tid == 0) { if (u_lo < u_hi) issue(u_lo, 0); if (u_lo + 1 < u_hi) issue(u_lo + 1, 1); }

    // weights as B fragments: b[s][nb][h] = (k = 16s + 8h + 2tg, +1 ; n = 8nb + g), W/255 rounded to bf16
    uint32_t wb[2][3][2];
    float bias0[3], bias1[3];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int nb = 0; nb < 3; ++nb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 16 * s + 8 * h + 2 * tg, n = 8 * nb + g;
                const float w0 = k < 27 ? a.w[k * kSC + n] * (1.f / 255.f) : 0.f;
                const float w1 = k + 1 < 27 ? a.w[(k + 1) * kSC + n] * (1.f / 255.f) : 0.f;
                wb[s][nb][h] = pack2(w0, w1);
            }
#pragma unroll
    for (int nb = 0; nb < 3; ++nb) { bias0[nb] = a.bias[8 * nb + 2 * tg]; bias1[nb] = a.bias[8 * nb + 2 * tg + 1]; }
    int koff[2][2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int e = 0; e < 2; ++e) koff[s][h][e] = stem_koff(16 * s + 8 * h + 2 * tg + e, G.W3);

    float ssum[6], ssq[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
    auto flush = [&](int t) {                          // CTA-uniform
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            float s = ssum[i], q = ssq[i];
            s += __shfl_xor_sync(0xffffffffu, s, 4); q += __shfl_xor_sync(0xffffffffu, q, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 8); q += __shfl_xor_sync(0xffffffffu, q, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16); q += __shfl_xor_sync(0xffffffffu, q, 16);
            if (g == 0) { const int c = 8 * (i >> 1) + 2 * tg + (i & 1); atomicAdd(&s_stat[2 * c], s); atomicAdd(&s_stat[2 * c + 1], q); }
            ssum[i] = 0.f; ssq[i] = 0.f;
        }
        __syncthreads();
        if (tid < kSC) {
            double2* dst = stat_slot(a.fst, kSC, (int)(blockIdx.x & (kStatCopies - 1)), t, tid);
            atomicAdd(&dst->x, (double)s_stat[2 * tid]); atomicAdd(&dst->y, (double)s_stat[2 * tid + 1]);
            s_stat[2 * tid] = 0.f; s_stat[2 * tid + 1] = 0.f;
        }
        __syncthreads();
    };

    int cur_t = -1;
    for (int u = u_lo, it = 0; u < u_hi; ++u, ++it) {
        const int buf = it & 1;
        const int f = u / a.nbands, band = u - f * a.nbands, t = f / G.B;
        const int y0 = band * a.HB, n = min(a.HB, G.Hs - y0), npx = n * G.Ws;
        if (t != cur_t) { if (cur_t >= 0 && a.training) flush(cur_t); cur_t = t; }
        mbar_wait(&full[buf], (it >> 1) & 1);
        if (tid == 0) bulk_store_wait_read();          // the previous band's store has left the staging tile
        __syncthreads();
        const uint8_t* ib = smem + L.img + (size_t)buf * L.img_stride;
        const int ngroups = (npx + 15) >> 4;
        for (int grp = warp; grp < ngroups; grp += kStemWarps) {
            const int q_lo = grp * 16 + g, q_hi = q_lo + 8;
            const uint8_t* p_lo = ib + (tbl[min(q_lo, npx - 1)] & 0xffffu);
            const uint8_t* p_hi = ib + (tbl[min(q_hi, npx - 1)] & 0xffffu);
            float acc[3][4];
#pragma unroll
            for (int nb = 0; nb < 3; ++nb) { acc[nb][0] = acc[nb][2] = bias0[nb]; acc[nb][1] = acc[nb][3] = bias1[nb]; }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint32_t af[4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    af[2 * h] = pack2((float)p_lo[koff[s][h][0]], (float)p_lo[koff[s][h][1]]);
                    af[2 * h + 1] = pack2((float)p_hi[koff[s][h][0]], (float)p_hi[koff[s][h][1]]);
                }
#pragma unroll
                for (int nb = 0; nb < 3; ++nb) mma16816(acc[nb], af, wb[s][nb][0], wb[s][nb][1]);
            }
#pragma unroll
            for (int nb = 0; nb < 3; ++nb) {
                const uint32_t v_lo = pack2(acc[nb][0], acc[nb][1]), v_hi = pack2(acc[nb][2], acc[nb][3]);
                if (q_lo < npx) {
                    *reinterpret_cast<uint32_t*>(tile + (size_t)q_lo * kSC + 8 * nb + 2 * tg) = v_lo;
                    const float2 v = unpack2(v_lo);
                    ssum[2 * nb] += v.x; ssq[2 * nb] = fmaf(v.x, v.x, ssq[2 * nb]);
                    ssum[2 * nb + 1] += v.y; ssq[2 * nb + 1] = fmaf(v.y, v.y, ssq[2 * nb + 1]);
                }
                if (q_hi < npx) {
                    *reinterpret_cast<uint32_t*>(tile + (size_t)q_hi * kSC + 8 * nb + 2 * tg) = v_hi;
                    const float2 v = unpack2(v_hi);
                    ssum[2 * nb] += v.x; ssq[2 * nb] = fmaf(v.x, v.x, ssq[2 * nb]);
                    ssum[2 * nb + 1] += v.y; ssq[2 * nb + 1] = fmaf(v.y, v.y, ssq[2 * nb + 1]);
                }
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(a.out + ((size_t)f * G.Hs + y0) * G.Ws * kSC, tile, (uint32_t)npx * kSC * 2);
            if (u + 2 < u_hi) issue(u + 2, buf);
        }
    }
    if (cur_t >= 0 && a.training) flush(cur_t);
    if (tid == 0) bulk_store_wait_all();
}

// ======================================================================================== max pool forward
struct PoolFwdArgs {
    const bf16* stem; const float2* aff;     // raw stem, (scale, shift) [4][24]
    StemGeom g;
    bf16* pool;                              // [kT*B][Hp*Wp][24] activated maxima
    uint8_t* idx;                            // [kT*B][Hp*Wp][24] window position (ky*3 + kx) of the first maximum
    int PB, nbands, units_per_cta;
};
struct PoolFwdSmem { int aff, tile, total; };
inline __host__ __device__ PoolFwdSmem pool_fwd_smem(int PB, int Ws) {
    PoolFwdSmem s; int off = 64;
    s.aff = off; off += kSC * 8;
    off = (off + 127) & ~127;
    s.tile = off; off += (2 * PB + 1) * Ws * kSC * 2;
    s.total = (off + 127) & ~127;
    return s;
}

__global__ void __launch_bounds__(kStemThreads, 4) pool_fwd_kernel(const PoolFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const StemGeom& G = a.g;
    const PoolFwdSmem L = pool_fwd_smem(a.PB, G.Ws);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    float2* s_aff = reinterpret_cast<float2*>(smem + L.aff);
    bf16* tile = reinterpret_cast<bf16*>(smem + L.tile);
    const int tid = threadIdx.x;
    const int nunits = kT * G.B * a.nbands;
    const int u_lo = blockIdx.x * a.units_per_cta, u_hi = min(nunits, u_lo + a.units_per_cta);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_fence_init(); }
    __syncthreads();
    auto rows_of = [&](int band, int& py0, int& py1, int& ya, int& yb) {
        py0 = band * a.PB; py1 = min(G.Hp, py0 + a.PB);
        ya = max(0, 2 * py0 - G.pad_t); yb = min(G.Hs, 2 * (py1 - 1) - G.pad_t + 3);
    };
    auto issue = [&](int u) {
        const int f = u / a.nbands, band = u - f * a.nbands;
        int py0, py1, ya, yb; rows_of(band, py0, py1, ya, yb);
        const uint32_t bytes = (uint32_t)(yb - ya) * G.Ws * kSC * 2;
        mbar_expect_tx(&full[0], bytes);
        bulk_g2s(tile, a.stem + ((size_t)f * G.Hs + ya) * G.Ws * kSC, bytes, &full[0]);
    };
    if (tid == 0 && u_lo < u_hi) issue(u_lo);
    // activation role: thread <-> (8-channel chunk ch, pixel lane); 255 threads busy
    const int ach = tid % 3, apl = tid / 3;
    int cur_t = -1;
    float2 c8[8];
    for (int u = u_lo, it = 0; u < u_hi; ++u, ++it) {
        const int f = u / a.nbands, band = u - f * a.nbands, t = f / G.B;
        int py0, py1, ya, yb; rows_of(band, py0, py1, ya, yb);
        if (t != cur_t) {
            __syncthreads();
            if (tid < kSC) s_aff[tid] = a.aff[(size_t)t * kSC + tid];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q] = s_aff[ach * 8 + q];
            cur_t = t;
        }
        mbar_wait(&full[0], it & 1);
        const int npx = (yb - ya) * G.Ws;
        if (tid < 255) {
            uint4* tv = reinterpret_cast<uint4*>(tile);
            for (int q = apl; q < npx; q += 85) tv[q * 3 + ach] = affine8(tv[q * 3 + ach], c8, true);
        }
        __syncthreads();
        const int nitem = (py1 - py0) * G.Wp * (kSC / 2);
        for (int i = tid; i < nitem; i += kStemThreads) {
            const int w = i / (kSC / 2), cp = i - w * (kSC / 2);
            const int wy = w / G.Wp, px = w - wy * G.Wp, py = py0 + wy;
            float b0 = -INFINITY, b1 = -INFINITY; int i0 = 0, i1 = 0;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int y = 2 * py - G.pad_t + ky;
                if (y < 0 || y >= G.Hs) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int x = 2 * px - G.pad_l + kx;
                    if (x < 0 || x >= G.Ws) continue;
                    const float2 v = unpack2(*reinterpret_cast<const uint32_t*>(tile + ((size_t)(y - ya) * G.Ws + x) * kSC + 2 * cp));
                    if (v.x > b0) { b0 = v.x; i0 = ky * 3 + kx; }
                    if (v.y > b1) { b1 = v.y; i1 = ky * 3 + kx; }
                }
            }
            const size_t o = ((size_t)f * G.Hp * G.Wp + (size_t)py * G.Wp + px) * kSC + 2 * cp;
            *reinterpret_cast<uint32_t*>(a.pool + o) = pack2(b0, b1);
            *reinterpret_cast<uint16_t*>(a.idx + o) = (uint16_t)(i0 | (i1 << 8));
        }
        __syncthreads();
        if (tid == 0 && u + 1 < u_hi) issue(u + 1);
    }
}

// ======================================================================================== stem backward
struct StemBwdArgs {
    const uint8_t* img; StemGeom g;
    const bf16* stem; const bf16* dpool; const uint8_t* idx;
    const float2* aff; const float2* bnp;    // [4][24]
    double* gacc;                            // [4][kGaccN], zeroed
    int HB, nbands, units_per_cta;
};
struct StemBwdSmem { int tbl, cst, red, img, raw, dz, dp, idx, total, nwr_max; };
inline __host__ __device__ StemBwdSmem stem_bwd_smem(int HB, int Ws, int Wp, int W3) {
    StemBwdSmem s; int off = 64;
    s.nwr_max = HB / 2 + 2;
    s.tbl = off; off += ((HB * Ws + 15) & ~15) * 4;
    s.cst = off; off += kSC * 16;
    s.red = off; off += kGaccN * 4;
    off = (off + 127) & ~127;
    s.img = off; off += ((2 * HB + 1) * W3 + 15 + 127) & ~127;
    s.raw = off; off += (HB * Ws * kSC * 2 + 127) & ~127;
    s.dz = off; off += (HB * Ws * kSC * 4 + 127) & ~127;      // fp32: a pixel can win several windows, the sum is not rounded per add
    s.dp = off; off += (s.nwr_max * Wp * kSC * 2 + 127) & ~127;
    s.idx = off; off += (s.nwr_max * Wp * kSC + 32 + 127) & ~127;
    s.total = off;
    return s;
}

__global__ void __launch_bounds__(kStemThreads, 2) stem_bwd_kernel(const StemBwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const StemGeom& G = a.g;
    const StemBwdSmem L = stem_bwd_smem(a.HB, G.Ws, G.Wp, G.W3);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint32_t* tbl = reinterpret_cast<uint32_t*>(smem + L.tbl);
    float4* s_cst = reinterpret_cast<float4*>(smem + L.cst);
    float* s_red = reinterpret_cast<float*>(smem + L.red);
    const uint8_t* ib = smem + L.img;
    const unsigned short* rawt = reinterpret_cast<const unsigned short*>(smem + L.raw);
    float* dzt = reinterpret_cast<float*>(smem + L.dz);
    const bf16* dpt = reinterpret_cast<const bf16*>(smem + L.dp);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const int nunits = kT * G.B * a.nbands;
    const int u_lo = blockIdx.x * a.units_per_cta, u_hi = min(nunits, u_lo + a.units_per_cta);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_fence_init(); }
    stem_fill_table(tbl, a.HB * G.Ws, G);
    for (int i = tid; i < kGaccN; i += kStemThreads) s_red[i] = 0.f;
    __syncthreads();

    const size_t frame_bytes = (size_t)G.H * G.W3;
    auto rows_of = [&](int band, int& y0, int& y1, int& pya, int& pyb) {
        y0 = band * a.HB; y1 = min(G.Hs, y0 + a.HB);
        pya = max(0, (y0 + G.pad_t - 1) / 2); pyb = min(G.Hp - 1, (y1 - 1 + G.pad_t) / 2);
    };
    auto idx_skip = [&](int f, int pya) { return (int)((((size_t)f * G.Hp + pya) * G.Wp * kSC) & 15); };
    auto issue = [&](int u) {                          // thread 0
        const int f = u / a.nbands, band = u - f * a.nbands, t = f / G.B, b = f - t * G.B;
        int y0, y1, pya, pyb; rows_of(band, y0, y1, pya, pyb);
        const int n = y1 - y0, nwr = pyb - pya + 1;
        const uint32_t b_img = (uint32_t)(((2 * n + 1) * G.W3 + 15) & ~15), b_raw = (uint32_t)n * G.Ws * kSC * 2;
        const uint32_t b_dp = (uint32_t)nwr * G.Wp * kSC * 2;
        const size_t io = ((size_t)f * G.Hp + pya) * G.Wp * kSC;
        const uint32_t b_idx = (uint32_t)(((io & 15) + (size_t)nwr * G.Wp * kSC + 15) & ~(size_t)15);
        mbar_expect_tx(&full[0], b_img + b_raw + b_dp + b_idx);
        bulk_g2s(smem + L.img, a.img + (size_t)(b * kT + t) * frame_bytes + (size_t)2 * y0 * G.W3, b_img, &full[0]);
        bulk_g2s(smem + L.raw, a.stem + ((size_t)f * G.Hs + y0) * G.Ws * kSC, b_raw, &full[0]);
        bulk_g2s(smem + L.dp, a.dpool + ((size_t)f * G.Hp + pya) * G.Wp * kSC, b_dp, &full[0]);
        bulk_g2s(smem + L.idx, a.idx + (io & ~(size_t)15), b_idx, &full[0]);
    };
    if (tid == 0 && u_lo < u_hi) issue(u_lo);

    // k-index rows of this thread's A fragments: g, g+8, g+16, g+24
    int koffT[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) koffT[j] = stem_koff(g + 8 * j, G.W3);
    float g1[2][3][4], g2[2][3][4], g0[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) { g1[m][j][e] = 0.f; g2[m][j][e] = 0.f; }
#pragma unroll
        for (int e = 0; e < 4; ++e) g0[m][e] = 0.f;
    }
    float s1[3] = {0.f, 0.f, 0.f}, s2[3] = {0.f, 0.f, 0.f};
    float4 cst[3];                                     // (scale, shift, inv_std, -mean * inv_std) of channels g, g+8, g+16

    auto flush = [&](int t) {                          // CTA-uniform: warps -> shared (float atomics) -> fp64 atomics
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = 16 * m + g + (e >> 1) * 8, c = 8 * j + 2 * tg + (e & 1);
                    atomicAdd(&s_red[k * kSC + c], g1[m][j][e]); atomicAdd(&s_red[32 * kSC + k * kSC + c], g2[m][j][e]);
                    g1[m][j][e] = 0.f; g2[m][j][e] = 0.f;
                }
            if (tg == 0) { atomicAdd(&s_red[64 * kSC + 16 * m + g], g0[m][0]); atomicAdd(&s_red[64 * kSC + 16 * m + g + 8], g0[m][2]); }
#pragma unroll
            for (int e = 0; e < 4; ++e) g0[m][e] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float x = s1[j], y = s2[j];
            x += __shfl_xor_sync(0xffffffffu, x, 1); y += __shfl_xor_sync(0xffffffffu, y, 1);
            x += __shfl_xor_sync(0xffffffffu, x, 2); y += __shfl_xor_sync(0xffffffffu, y, 2);
            if (tg == 0) { atomicAdd(&s_red[64 * kSC + 32 + g + 8 * j], x); atomicAdd(&s_red[64 * kSC + 32 + kSC + g + 8 * j], y); }
            s1[j] = 0.f; s2[j] = 0.f;
        }
        __syncthreads();
        for (int i = tid; i < kGaccN; i += kStemThreads) {
            const float v = s_red[i];
            if (v != 0.f) atomicAdd(a.gacc + (size_t)t * kGaccN + i, (double)v);
            s_red[i] = 0.f;
        }
        __syncthreads();
    };

    int cur_t = -1;
    for (int u = u_lo, it = 0; u < u_hi; ++u, ++it) {
        const int f = u / a.nbands, band = u - f * a.nbands, t = f / G.B;
        int y0, y1, pya, pyb; rows_of(band, y0, y1, pya, pyb);
        const int n = y1 - y0, npx = n * G.Ws, nwr = pyb - pya + 1;
        if (t != cur_t) {
            if (cur_t >= 0) flush(cur_t);
            __syncthreads();
            if (tid < kSC) {
                const float2 af = a.aff[(size_t)t * kSC + tid], bp = a.bnp[(size_t)t * kSC + tid];
                s_cst[tid] = make_float4(af.x, af.y, bp.y, -bp.x * bp.y);
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 3; ++j) cst[j] = s_cst[g + 8 * j];
            cur_t = t;
        }
        // ---- zero the dz tile (overlaps the loads)
        for (int i = tid; i < (npx * kSC * 4 + 15) / 16; i += kStemThreads) reinterpret_cast<uint4*>(dzt)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        mbar_wait(&full[0], it & 1);
        // ---- max-pool backward: every window adds its gradient to the stored winner (if that pixel is in this band)
        {
            const uint8_t* it8 = smem + L.idx + idx_skip(f, pya);
            const int nitem = nwr * G.Wp * (kSC / 2);
            for (int i = tid; i < nitem; i += kStemThreads) {
                const int w = i / (kSC / 2), cp = i - w * (kSC / 2);
                const int wy = w / G.Wp, px = w - wy * G.Wp, py = pya + wy;
                const uint32_t code = *reinterpret_cast<const uint16_t*>(it8 + (size_t)w * kSC + 2 * cp);
                const uint32_t d = *reinterpret_cast<const uint32_t*>(dpt + (size_t)w * kSC + 2 * cp);
                const int c0 = code & 0xff, c1 = code >> 8;
                const int ya0 = 2 * py - G.pad_t + c0 / 3 - y0, xa0 = 2 * px - G.pad_l + c0 % 3;
                const int ya1 = 2 * py - G.pad_t + c1 / 3 - y0, xa1 = 2 * px - G.pad_l + c1 % 3;
                const float2 dv = unpack2(d);
                if (ya0 >= 0 && ya0 < n) atomicAdd(dzt + ((size_t)ya0 * G.Ws + xa0) * kSC + 2 * cp, dv.x);
                if (ya1 >= 0 && ya1 < n) atomicAdd(dzt + ((size_t)ya1 * G.Ws + xa1) * kSC + 2 * cp + 1, dv.y);
            }
        }
        __syncthreads();
        // ---- per 16-pixel group: ReLU6 mask, sums, and the three tensor-core products
        const int ngroups = (npx + 15) >> 4;
        for (int grp = warp; grp < ngroups; grp += kStemWarps) {
            uint32_t pa[4][2];          // [k row j][pixel pair h]: centred pixel values of (px 2h, 2h+1) for k = g + 8j
            uint32_t bz[3][2], bx[3][2], ones[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float pv[2][4];
                float dzv[2][3], xhv[2][3];
                bool valid[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int q = grp * 16 + 8 * h + 2 * tg + e;
                    valid[e] = q < npx;
                    const int qc = valid[e] ? q : npx - 1;
                    const uint8_t* pp = ib + (tbl[qc] & 0xffffu);
#pragma unroll
                    for (int j = 0; j < 4; ++j) pv[e][j] = (float)((int)pp[koffT[j]] - 128);
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float raw = __uint_as_float((uint32_t)rawt[(size_t)qc * kSC + g + 8 * j] << 16);
                        const float dzr = dzt[(size_t)qc * kSC + g + 8 * j];
                        const float uu = fmaf(raw, cst[j].x, cst[j].y);
                        const float dz = (valid[e] && uu > 0.f && uu < 6.f) ? dzr : 0.f;
                        const float xh = valid[e] ? fmaf(raw, cst[j].z, cst[j].w) : 0.f;
                        dzv[e][j] = dz; xhv[e][j] = xh;
                        s1[j] += dz; s2[j] = fmaf(dz, xh, s2[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) pa[j][h] = pack2(pv[0][j], pv[1][j]);
#pragma unroll
                for (int j = 0; j < 3; ++j) { bz[j][h] = pack2(dzv[0][j], dzv[1][j]); bx[j][h] = pack2(xhv[0][j], xhv[1][j]); }
                ones[h] = pack2(valid[0] ? 1.f : 0.f, valid[1] ? 1.f : 0.f);
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const uint32_t af[4] = {pa[2 * m][0], pa[2 * m + 1][0], pa[2 * m][1], pa[2 * m + 1][1]};
#pragma unroll
                for (int j = 0; j < 3; ++j) { mma16816(g1[m][j], af, bz[j][0], bz[j][1]); mma16816(g2[m][j], af, bx[j][0], bx[j][1]); }
                mma16816(g0[m], af, ones[0], ones[1]);
            }
        }
        __syncthreads();
        if (tid == 0 && u + 1 < u_hi) issue(u + 1);
    }
    if (cur_t >= 0) flush(cur_t);
}

// dW, dgamma, dbeta from the per-slice sums (one CTA)
struct StemFinArgs { const double* gacc; const float2* aff; float* dw; float* dgamma; float* dbeta; double n; };
__global__ void __launch_bounds__(kStemThreads) stem_bwd_finish_kernel(const StemFinArgs a) {
    for (int i = threadIdx.x; i < 27 * kSC; i += kStemThreads) {
        const int k = i / kSC, c = i - k * kSC;
        double acc = 0.0;
        for (int t = 0; t < kT; ++t) {
            const double* G = a.gacc + (size_t)t * kGaccN;
            const double S1 = G[64 * kSC + 32 + c], S2 = G[64 * kSC + 32 + kSC + c];
            acc += (double)a.aff[(size_t)t * kSC + c].x * (G[k * kSC + c] - S1 / a.n * G[64 * kSC + k] - S2 / a.n * G[32 * kSC + k * kSC + c]);
        }
        a.dw[i] = (float)(acc / 255.0);
    }
    for (int c = threadIdx.x; c < kSC; c += kStemThreads) {
        double gs = 0.0, bs = 0.0;
        for (int t = 0; t < kT; ++t) { const double* G = a.gacc + (size_t)t * kGaccN; bs += G[64 * kSC + 32 + c]; gs += G[64 * kSC + 32 + kSC + c]; }
        a.dgamma[c] = (float)gs; a.dbeta[c] = (float)bs;
    }
}

}  // namespace v2
}  // namespace cdra
#endif
