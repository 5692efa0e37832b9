// v2 tower: backward kernels -- what tape.gradient(loss, dynamics.trainable_variables) (core/carla_agent.py:361-365,
// 440-444) computes through the pointwise / depthwise layers, the channel shuffle and the global average pool.
//
// Gradient tensors hold d loss / d (activated value) in the layout of the forward tensor.  A layer's own BatchNorm
// (+ReLU6) backward is applied when its output gradient is loaded:
//     dR = scale * (dz - S1/n - xhat * S2/n),   dz = dA * [0 < z < 6],   xhat = (R - mean) * inv_std
// where S1 = sum dz, S2 = sum dz * xhat over the slice.  Those sums are produced by whichever kernel WRITES the final
// gradient of a tensor (it has the raw tensor in shared memory anyway), so no separate statistics pass exists.
// Biases that feed a training-mode BatchNorm have an analytically zero gradient: it is written as 0.
#pragma once
#ifndef CDRA_EMU
#include "v2_pw.cuh"
#include "v2_dw.cuh"
#include "v2_umma.cuh"

namespace cdra {
namespace v2 {

// per-column constants of the BatchNorm(+ReLU6) backward:  dR = fma(dz, x, fma(raw, w, z)),  mask from fma(raw, x, y)
CDRA_DEV float4 bnbwd_consts(float2 aff, float2 bnp, double2 bsum, double inv_n) {
    const float k1 = (float)(bsum.x * inv_n), k2 = (float)(bsum.y * inv_n);
    float4 c;
    c.x = aff.x;                                   // scale
    c.y = aff.y;                                   // shift
    c.z = aff.x * (k2 * bnp.x * bnp.y - k1);       // A0 = scale * (k2 * mean * inv - k1)
    c.w = -aff.x * k2 * bnp.y;                     // B1 = -scale * k2 * inv
    return c;
}
CDRA_DEV float bnbwd_apply(float dA, float raw, const float4& c, bool clamp) {
    const float z = fmaf(raw, c.x, c.y);
    const float dz = (!clamp || (z > 0.f && z < 6.f)) ? dA : 0.f;
    return fmaf(dz, c.x, fmaf(raw, c.w, c.z));
}
// per-slot constants for the sums of a freshly written gradient: (scale, shift, inv, -mean*inv)
CDRA_DEV float4 sum_consts(const float2* aff, const float2* bnp, size_t idx) {
    if (!aff) return make_float4(1.f, 0.f, 0.f, 0.f);
    const float2 a = aff[idx], b = bnp[idx];
    return make_float4(a.x, a.y, b.y, -b.x * b.y);
}
CDRA_DEV void sum_accum(float dA, float raw, const float4& c, bool clamp, float& s1, float& s2) {
    const float z = fmaf(raw, c.x, c.y);
    const float dz = (!clamp || (z > 0.f && z < 6.f)) ? dA : 0.f;
    s1 += dz;
    s2 = fmaf(dz, fmaf(raw, c.z, c.w), s2);
}

struct PwBwdArgs {
    const PwDesc* d;
    int Rt;
    const bf16* out[2]; const bf16* dout[2]; int cpo; Tables tb[2];   // the layer's output plane(s): raw, gradient, tables
    int out_clamp;
    // pass-through half of a stride-1 unit: d x1[slot(2i+p)] = d out_p[copy_dst0 + i]
    const bf16* x1; bf16* dx1; int x1cp; SlotMap x1map; const float2* x1aff; const float2* x1bnp; double2* x1bsum;
    int x1clamp, ncopy, copy_dst0;
    int tiles_per_cta, nbuf;
    int direct;                 // 1: d out / out are read straight from HBM by the transform (very wide layers), not TMA-staged
    int ntiles_k[kMaxSrc];      // dgrad: K tiles per source
    // wgrad tiling
    int kt_tiles, nt_tiles;
    unsigned* counter;
    // dR = BatchNorm(+ReLU6) backward of d out, [4*Rt][NPall] bf16: written once by the data-gradient kernel (it builds
    // the tile anyway), consumed by the tcgen05 weight-gradient kernel, which then needs no transform of its own
    bf16* dr;
    GBlock blk[kGMaxBlk]; int nblk;   // v4_pwg.cuh: blockIdx.y = M block
    int timeline;                     // record the role timeline of block 0 (cdra_debug_timeline)
};

struct PwDgradSmem { int colc, srcc, x1c, stat, w, raw, dr, st, st2, total, raw_stride, ldr, ldw, lds, lds2; };
inline __host__ __device__ PwDgradSmem pw_dgrad_smem(int R, int KT, int NPall, int nplanes, int cpo, int src_cp, int x1cp, int nbuf, int direct, int nt = 256) {
    PwDgradSmem s;
    s.ldr = pad_ld(NPall); s.ldw = pad_ld(NPall); s.lds = pad_ld(KT); s.lds2 = pad_ld(x1cp);
    int off = 64;
    s.colc = off; off += NPall * 16;
    s.srcc = off; off += KT * 16;
    s.x1c = off; off += x1cp * 16;
    s.stat = off; off += (KT + x1cp) * 8;
    off = (off + 127) & ~127;
    s.w = off; off += KT * s.ldw * 2;
    off = (off + 127) & ~127;
    s.raw_stride = (R * ((direct ? 0 : 2 * nplanes * cpo) + src_cp + x1cp) * 2 + 127) & ~127;
    s.raw = off; off += nbuf * s.raw_stride;
    { const int dr_bytes = R * s.ldr * 2, scr = nt * 128; s.dr = off; off += dr_bytes > scr ? dr_bytes : scr; }     // doubles as the flush scratch (2 x nt*8 float2)
    off = (off + 127) & ~127;
    s.st = off; off += R * s.lds * 2;
    off = (off + 127) & ~127;
    s.st2 = off; off += x1cp ? R * s.lds2 * 2 : 0;
    s.total = off;
    return s;
}

// ======================================================================================== data gradient
// d src[r][kk] = sum_j dR[r][j] * W[kk][j]  for the K tile (source, k0..k0+kw) of this CTA; epilogue: store (or add to)
// the source's gradient tensor and accumulate its BatchNorm-backward sums.
template <int R, int WM, int WN, int MT, int NBW>
__global__ void __launch_bounds__(WM * WN * 32, (WM * WN == 8 ? 2 : 1)) pw_dgrad_kernel(const PwBwdArgs a) {
    constexpr int KT = WN * NBW * 8, NT = WM * WN * 32;     // 8 warps (two CTAs per SM) or 16 warps (one CTA per SM)
    extern __shared__ __align__(128) unsigned char smem[];
    const PwDesc& d = *a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int NP = d.NPall, gwp = d.cols.gwp, nplanes = d.cols.nplanes;
    // ---- this CTA's K tile
    int si = 0, ky = blockIdx.y, koff = 0;
    while (ky >= a.ntiles_k[si]) { ky -= a.ntiles_k[si]; koff += d.src[si].cp; ++si; }
    const PwSrc& S = d.src[si];
    const int k0 = ky * KT, kw = min(KT, S.cp - k0);
    const bool do_x1 = a.x1 != nullptr && blockIdx.y == 0;
    const int x1cp = do_x1 ? a.x1cp : 0;
    const PwDgradSmem L = pw_dgrad_smem(R, KT, NP, nplanes, a.cpo, S.cp, a.x1 ? a.x1cp : 0, a.nbuf, a.direct, NT);
    const int NP16 = (NP + 15) & ~15;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    float4* s_colc = reinterpret_cast<float4*>(smem + L.colc);
    float4* s_srcc = reinterpret_cast<float4*>(smem + L.srcc);
    float4* s_x1c = reinterpret_cast<float4*>(smem + L.x1c);
    float* s_stat = reinterpret_cast<float*>(smem + L.stat);
    bf16* Ws = reinterpret_cast<bf16*>(smem + L.w);
    unsigned char* raw = smem + L.raw;
    bf16* Dr = reinterpret_cast<bf16*>(smem + L.dr);
    bf16* St = reinterpret_cast<bf16*>(smem + L.st);
    bf16* St2 = reinterpret_cast<bf16*>(smem + L.st2);
    const int ldr = L.ldr, ldw = L.ldw, lds = L.lds, lds2 = L.lds2;
    // raw buffer layout (rows R each): [dout p0][dout p1][out p0][out p1][src][x1]
    const int plane_bytes = a.cpo * 2;
    const int o_dout = 0, o_out = nplanes * plane_bytes, o_src = a.direct ? 0 : 2 * nplanes * plane_bytes, o_x1 = o_src + S.cp * 2;

    const int tps = (a.Rt + R - 1) / R, ntile = kT * tps;
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(ntile, tile_lo + a.tiles_per_cta);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    for (int i = tid; i < KT * (ldw / 8); i += NT) {
        const int k = i / (ldw / 8), c = i - k * (ldw / 8);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (k < kw && c < NP / 8) v = *reinterpret_cast<const uint4*>(d.wb + (size_t)(koff + k0 + k) * NP + c * 8);
        *reinterpret_cast<uint4*>(Ws + (size_t)k * ldw + c * 8) = v;
    }
    for (int i = tid; i < R * ldr / 2; i += NT) reinterpret_cast<uint32_t*>(Dr)[i] = 0u;
    for (int i = tid; i < R * lds / 2; i += NT) reinterpret_cast<uint32_t*>(St)[i] = 0u;
    if (do_x1) for (int i = tid; i < R * lds2 / 2; i += NT) reinterpret_cast<uint32_t*>(St2)[i] = 0u;
    for (int i = tid; i < (KT + x1cp) * 2; i += NT) s_stat[i] = 0.f;
    __syncthreads();

    auto issue = [&](int tile, int buf) {
        const int t = tile / tps, r0 = (tile - t * tps) * R, rows = min(R, a.Rt - r0);
        unsigned char* dst = raw + (size_t)buf * L.raw_stride;
        const size_t row = (size_t)t * a.Rt + r0;
        uint32_t bytes = rows * ((a.direct ? 0 : 2 * nplanes * plane_bytes) + S.cp * 2 + x1cp * 2);
        mbar_expect_tx(&full[buf], bytes);
        if (!a.direct) for (int p = 0; p < nplanes; ++p) {
            bulk_g2s(dst + (size_t)R * (o_dout + p * plane_bytes), a.dout[p] + row * a.cpo, rows * plane_bytes, &full[buf]);
            bulk_g2s(dst + (size_t)R * (o_out + p * plane_bytes), a.out[p] + row * a.cpo, rows * plane_bytes, &full[buf]);
        }
        bulk_g2s(dst + (size_t)R * o_src, S.data + row * S.cp, rows * S.cp * 2, &full[buf]);
        if (do_x1) bulk_g2s(dst + (size_t)R * o_x1, a.x1 + row * a.x1cp, rows * x1cp * 2, &full[buf]);
    };
    pdl_wait();
    if (tid == 0) for (int b = 0; b < a.nbuf; ++b) if (tile_lo + b < tile_hi) issue(tile_lo + b, b);

    const int wm = warp % WM, wn = warp / WM, g = lane >> 2, tg = lane & 3;
    // roles
    const int nq = nplanes * (gwp >> 3), tq = tid % nq, trl = tid / nq, tnrl = NT / nq;            // dR transform
    const int tp = tq / (gwp >> 3), tc = (tq - tp * (gwp >> 3)) * 8;
    const int nv = kw >> 3, vq = tid % nv, vrl = tid / nv, vnrl = NT / nv;                          // store pass
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    const int nx = do_x1 ? (x1cp >> 3) : 1, xq = tid % nx, xrl = tid / nx, xnrl = NT / nx;          // pass-through store pass
    float x1s1[8], x1s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x1s1[i] = 0.f; x1s2[i] = 0.f; }
    int cs_src = -1, cs_rl = 0, cs_nrl = 1, cs_slot = 0;                                            // pass-through gather
    if (do_x1) {
        cs_nrl = NT / x1cp; cs_rl = tid / x1cp; cs_slot = tid % x1cp;
        if (cs_rl < cs_nrl) {
            const int l = slot_logical(a.x1map, cs_slot);
            if (l >= 0 && (l >> 1) < a.ncopy) cs_src = (l & 1) * plane_bytes / 2 * R + a.copy_dst0 + (l >> 1);   // element offset inside the dout region
            else cs_src = -2;                                                                      // padding -> zero gradient
        }
    }
    const bool sclamp = S.clamp != 0, oclamp = a.out_clamp != 0, xclamp = a.x1clamp != 0;
    const double inv_n = 1.0 / (double)a.Rt;

    // sums: registers -> scratch (the idle dR tile) -> one thread per column -> fp64 atomics (no shared-memory float atomics)
    auto flush = [&](int t) {
        __syncthreads();
        float2* scr = reinterpret_cast<float2*>(Dr);              // [vnrl][kw] then [xnrl][x1cp]; <= 2 * NT*8 entries
        if (vrl < vnrl) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { scr[vrl * kw + vq * 8 + i] = make_float2(s1[i], s2[i]); s1[i] = 0.f; s2[i] = 0.f; }
        }
        float2* scr2 = scr + NT * 8;
        if (do_x1 && xrl < xnrl) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { scr2[xrl * x1cp + xq * 8 + i] = make_float2(x1s1[i], x1s2[i]); x1s1[i] = 0.f; x1s2[i] = 0.f; }
        }
        __syncthreads();
        for (int i = tid; i < kw; i += NT) {
            const int s = k0 + i;
            if (s >= S.sum_lo && s < S.sum_hi) {
                float a1 = 0.f, a2 = 0.f;
                for (int l = 0; l < vnrl; ++l) { const float2 v = scr[l * kw + i]; a1 += v.x; a2 += v.y; }
                double2* dst = S.bsum + (size_t)t * S.cp + s;
                atomicAdd(&dst->x, (double)a1); atomicAdd(&dst->y, (double)a2);
            }
        }
        if (do_x1) for (int i = tid; i < x1cp; i += NT) {
            float a1 = 0.f, a2 = 0.f;
            for (int l = 0; l < xnrl; ++l) { const float2 v = scr2[l * x1cp + i]; a1 += v.x; a2 += v.y; }
            double2* dst = a.x1bsum + (size_t)t * x1cp + i;
            atomicAdd(&dst->x, (double)a1); atomicAdd(&dst->y, (double)a2);
        }
        __syncthreads();
        for (int i = tid; i < R * ldr / 2; i += NT) reinterpret_cast<uint32_t*>(Dr)[i] = 0u;      // K padding columns back to zero
        __syncthreads();
    };

    int cur_t = -1;
    for (int tile = tile_lo, it = 0; tile < tile_hi; ++tile, ++it) {
        const int buf = it % a.nbuf;
        const int t = tile / tps, r0 = (tile - t * tps) * R, rows = min(R, a.Rt - r0);
        if (t != cur_t) {
            if (cur_t >= 0) flush(cur_t);
            __syncthreads();
            for (int j = tid; j < NP; j += NT) {
                int p, s, l, n;
                float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pw_col(d, j, p, s, l, n)) {
                    const size_t idx = (size_t)t * a.cpo + s;
                    c = bnbwd_consts(a.tb[p].aff[idx], a.tb[p].bnp[idx], a.tb[p].bsum[idx], inv_n);
                }
                s_colc[tcol(j, NP >> 3)] = c;
            }
            for (int i = tid; i < kw; i += NT) s_srcc[tcol(i, KT >> 3)] = sum_consts(S.aff, S.bnp, (size_t)t * S.cp + k0 + i);
            if (do_x1) for (int i = tid; i < x1cp; i += NT) s_x1c[tcol(i, x1cp >> 3)] = sum_consts(a.x1aff, a.x1bnp, (size_t)t * x1cp + i);
            cur_t = t;
            __syncthreads();
        }
        mbar_wait(&full[buf], (it / a.nbuf) & 1);
        const unsigned char* rb = raw + (size_t)buf * L.raw_stride;
        // ---- dR tile from (d out, out): BatchNorm + ReLU6 backward on load
        if (trl < tnrl) {
            float4 c8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q] = s_colc[q * (NP >> 3) + ((tp * gwp + tc) >> 3)];
            const uint4* dv = reinterpret_cast<const uint4*>(rb + (size_t)R * (o_dout + tp * plane_bytes));
            const uint4* ov = reinterpret_cast<const uint4*>(rb + (size_t)R * (o_out + tp * plane_bytes));
            if (a.direct) {
                dv = reinterpret_cast<const uint4*>(a.dout[tp] + ((size_t)t * a.Rt + r0) * a.cpo);
                ov = reinterpret_cast<const uint4*>(a.out[tp] + ((size_t)t * a.Rt + r0) * a.cpo);
            }
            const int nch = a.cpo >> 3, ch = tc >> 3;
            for (int r = trl; r < rows; r += tnrl) {
                uint4 dvv = dv[r * nch + ch]; const uint4 ovv = ov[r * nch + ch];
                uint32_t* dw = reinterpret_cast<uint32_t*>(&dvv); const uint32_t* ow = reinterpret_cast<const uint32_t*>(&ovv);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 dd = unpack2(dw[i]), oo = unpack2(ow[i]);
                    dw[i] = pack2(bnbwd_apply(dd.x, oo.x, c8[2 * i], oclamp), bnbwd_apply(dd.y, oo.y, c8[2 * i + 1], oclamp));
                }
                *reinterpret_cast<uint4*>(Dr + (size_t)r * ldr + tp * gwp + tc) = dvv;
            }
        }
        // rows beyond `rows` of a partial tile keep stale (finite) values: their products are never stored
        __syncthreads();
        if (a.dr != nullptr && blockIdx.y == 0) {     // hand the finished dR tile to the weight-gradient kernel
            const int nch = NP >> 3;
            bf16* drow = a.dr + ((size_t)t * a.Rt + r0) * NP;
            for (int i = tid; i < rows * nch; i += NT) {
                const int r = i / nch, c = i - r * nch;
                *reinterpret_cast<uint4*>(drow + (size_t)r * NP + c * 8) = *reinterpret_cast<const uint4*>(Dr + (size_t)r * ldr + c * 8);
            }
        }
        // ---- MMA: d src tile [R x KT] = dR [R x NP] * Wb^T
        float acc[MT][NBW][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) { acc[mt][nb][0] = acc[mt][nb][1] = acc[mt][nb][2] = acc[mt][nb][3] = 0.f; }
        {
            const int arow = wm * MT * 16 + (lane & 15), acol = (lane >> 4) * 8;
            const int mi = lane >> 3;
            const int brow = wn * NBW * 8 + (mi >> 1) * 8 + (lane & 7), bcol = (mi & 1) * 8;
            for (int ks = 0; ks < NP16; ks += 16) {
                uint32_t af[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) ldsm4(af[mt], Dr + (size_t)(arow + mt * 16) * ldr + ks + acol);
#pragma unroll
                for (int nb2 = 0; nb2 < NBW / 2; ++nb2) {
                    uint32_t bfr[4];
                    ldsm4(bfr, Ws + (size_t)(brow + nb2 * 16) * ldw + ks + bcol);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma16816(acc[mt][2 * nb2], af[mt], bfr[0], bfr[1]);
                        mma16816(acc[mt][2 * nb2 + 1], af[mt], bfr[2], bfr[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) {
            const int jl = wn * NBW * 8 + nb * 8 + 2 * tg;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int r = wm * MT * 16 + mt * 16 + g;
                *reinterpret_cast<uint32_t*>(St + (size_t)r * lds + jl) = pack2(acc[mt][nb][0], acc[mt][nb][1]);
                *reinterpret_cast<uint32_t*>(St + (size_t)(r + 8) * lds + jl) = pack2(acc[mt][nb][2], acc[mt][nb][3]);
            }
        }
        // ---- pass-through half: gather d x1 from the output planes' gradient (bit-exact)
        if (cs_src != -1) {
            const bf16* dreg = reinterpret_cast<const bf16*>(rb + (size_t)R * o_dout);
            if (cs_src >= 0) {
                const int pl = cs_src / (plane_bytes / 2 * R), col = cs_src - pl * (plane_bytes / 2 * R);
                const bf16* srcp = dreg + (size_t)pl * R * a.cpo + col;
                for (int r = cs_rl; r < rows; r += cs_nrl) St2[(size_t)r * lds2 + cs_slot] = srcp[(size_t)r * a.cpo];
            }
        }
        __syncthreads();
        // ---- store d src (+ existing partial gradient), accumulate its BatchNorm-backward sums
        if (vrl < vnrl) {
            float4 c8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q] = s_srcc[q * (KT >> 3) + vq];
            bf16* grow = S.grad + ((size_t)t * a.Rt + r0) * S.cp + k0 + vq * 8;
            const uint4* rv = reinterpret_cast<const uint4*>(rb + (size_t)R * o_src);
            const int nch = S.cp >> 3, ch = (k0 >> 3) + vq;
            for (int r = vrl; r < rows; r += vnrl) {
                uint4 v = *reinterpret_cast<const uint4*>(St + (size_t)r * lds + vq * 8);
                uint32_t* w = reinterpret_cast<uint32_t*>(&v);
                if (S.accumulate) {
                    const uint4 e = *reinterpret_cast<const uint4*>(grow + (size_t)r * S.cp);
                    const uint32_t* ew = reinterpret_cast<const uint32_t*>(&e);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const float2 x = unpack2(w[i]), y = unpack2(ew[i]); w[i] = pack2(x.x + y.x, x.y + y.y); }
                }
                *reinterpret_cast<uint4*>(grow + (size_t)r * S.cp) = v;
                const uint4 rw = rv[r * nch + ch];
                const uint32_t* rww = reinterpret_cast<const uint32_t*>(&rw);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 x = unpack2(w[i]), y = unpack2(rww[i]);
                    sum_accum(x.x, y.x, c8[2 * i], sclamp, s1[2 * i], s2[2 * i]);
                    sum_accum(x.y, y.y, c8[2 * i + 1], sclamp, s1[2 * i + 1], s2[2 * i + 1]);
                }
            }
        }
        if (do_x1 && xrl < xnrl) {
            float4 c8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q] = s_x1c[q * (x1cp >> 3) + xq];
            bf16* grow = a.dx1 + ((size_t)t * a.Rt + r0) * x1cp + xq * 8;
            const uint4* rv = reinterpret_cast<const uint4*>(rb + (size_t)R * o_x1);
            const int nch = x1cp >> 3;
            for (int r = xrl; r < rows; r += xnrl) {
                const uint4 v = *reinterpret_cast<const uint4*>(St2 + (size_t)r * lds2 + xq * 8);
                *reinterpret_cast<uint4*>(grow + (size_t)r * x1cp) = v;
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
                const uint4 rw = rv[r * nch + xq];
                const uint32_t* rww = reinterpret_cast<const uint32_t*>(&rw);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 x = unpack2(w[i]), y = unpack2(rww[i]);
                    sum_accum(x.x, y.x, c8[2 * i], xclamp, x1s1[2 * i], x1s2[2 * i]);
                    sum_accum(x.y, y.y, c8[2 * i + 1], xclamp, x1s1[2 * i + 1], x1s2[2 * i + 1]);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && tile + a.nbuf < tile_hi) issue(tile + a.nbuf, buf);
    }
    if (cur_t >= 0) flush(cur_t);
}

// ======================================================================================== weight gradient
// dW[kk][j] += sum_r act(src)[r][kk] * dR[r][j] over this CTA's rows, for its (K tile, N tile); fp32 atomics into the
// (pre-zeroed) gradient arena in the reference's logical layout.  BN gamma/beta gradients come from the final sums.
constexpr int kWgK = 64, kWgR = 64;
struct PwWgradSmem { int colc, aff, raw, xs, rs, total, raw_stride, ldx, ldr; };
inline __host__ __device__ PwWgradSmem pw_wgrad_smem(int NTW, int cpo, int src_cp, int nbuf, int direct, int dr_cols = 0) {
    PwWgradSmem s;
    s.ldx = pad_ld(kWgK); s.ldr = pad_ld(NTW);
    int off = 64;
    s.colc = off; off += NTW * 16;
    s.aff = off; off += kWgK * 8;
    off = (off + 127) & ~127;
    s.raw_stride = (kWgR * ((dr_cols ? dr_cols : (direct ? 0 : 2 * cpo)) + src_cp) * 2 + 127) & ~127;    // dr_cols: finished dR rows instead of (d out, out)
    s.raw = off; off += nbuf * s.raw_stride;
    s.xs = off; off += kWgR * s.ldx * 2;
    off = (off + 127) & ~127;
    s.rs = off; off += kWgR * s.ldr * 2;
    s.total = off;
    return s;
}

template <int NBW>          // n-blocks per warp: N tile = 2 * NBW * 8 columns
__global__ void __launch_bounds__(256, 2) pw_wgrad_kernel(const PwBwdArgs a) {
    constexpr int NTW = 2 * NBW * 8;
    extern __shared__ __align__(128) unsigned char smem[];
    const PwDesc& d = *a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int NP = d.NPall, gwp = d.cols.gwp;
    // ---- tile: K tile (source si, k0) x N tile (plane pn, slot n0)
    const int kti = blockIdx.y / a.nt_tiles, nti = blockIdx.y - kti * a.nt_tiles;
    int si = 0, ky = kti, koff = 0;
    while (ky >= a.ntiles_k[si]) { ky -= a.ntiles_k[si]; koff += d.src[si].cp; ++si; }
    const PwSrc& S = d.src[si];
    const int k0 = ky * kWgK, kw = min(kWgK, S.cp - k0);
    const int ntp = (gwp + NTW - 1) / NTW;                 // N tiles per plane
    const int pn = nti / ntp, n0 = (nti - pn * ntp) * NTW, nw = min(NTW, gwp - n0);
    const bool have_dr = a.dr != nullptr;              // the data-gradient kernel left the finished dR matrix [4*Rt][NP]
    const PwWgradSmem L = pw_wgrad_smem(NTW, a.cpo, S.cp, a.nbuf, a.direct, have_dr ? NP : 0);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    float4* s_colc = reinterpret_cast<float4*>(smem + L.colc);
    float2* s_aff = reinterpret_cast<float2*>(smem + L.aff);
    unsigned char* raw = smem + L.raw;
    bf16* Xs = reinterpret_cast<bf16*>(smem + L.xs);
    bf16* Rs = reinterpret_cast<bf16*>(smem + L.rs);
    const int ldx = L.ldx, ldr = L.ldr;
    const int plane_bytes = a.cpo * 2;
    const int o_dout = 0, o_out = plane_bytes, o_src = have_dr ? NP * 2 : (a.direct ? 0 : 2 * plane_bytes);

    const int tps = (a.Rt + kWgR - 1) / kWgR, ntile = kT * tps;
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(ntile, tile_lo + a.tiles_per_cta);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    for (int i = tid; i < kWgR * ldx / 2; i += 256) reinterpret_cast<uint32_t*>(Xs)[i] = 0u;
    for (int i = tid; i < kWgR * ldr / 2; i += 256) reinterpret_cast<uint32_t*>(Rs)[i] = 0u;
    __syncthreads();
    auto issue = [&](int tile, int buf) {
        const int t = tile / tps, r0 = (tile - t * tps) * kWgR, rows = min(kWgR, a.Rt - r0);
        unsigned char* dst = raw + (size_t)buf * L.raw_stride;
        const size_t row = (size_t)t * a.Rt + r0;
        mbar_expect_tx(&full[buf], rows * ((have_dr ? NP * 2 : (a.direct ? 0 : 2 * plane_bytes)) + S.cp * 2));
        if (have_dr) bulk_g2s(dst, a.dr + row * NP, rows * NP * 2, &full[buf]);
        else if (!a.direct) {
            bulk_g2s(dst + (size_t)kWgR * o_dout, a.dout[pn] + row * a.cpo, rows * plane_bytes, &full[buf]);
            bulk_g2s(dst + (size_t)kWgR * o_out, a.out[pn] + row * a.cpo, rows * plane_bytes, &full[buf]);
        }
        bulk_g2s(dst + (size_t)kWgR * o_src, S.data + row * S.cp, rows * S.cp * 2, &full[buf]);
    };
    pdl_wait();
    if (tid == 0) for (int b = 0; b < a.nbuf; ++b) if (tile_lo + b < tile_hi) issue(tile_lo + b, b);

    const int kg = warp & 3, nh = warp >> 2;               // warp: k rows kg*16.., n columns nh*NBW*8..
    float acc[NBW][4];
#pragma unroll
    for (int nb = 0; nb < NBW; ++nb) { acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f; }
    const int nqr = nw >> 3, rq = tid % nqr, rrl = tid / nqr, rnrl = 256 / nqr;          // dR transform role
    const int nqx = kw >> 3, xq = tid % nqx, xrl = tid / nqx, xnrl = 256 / nqx;          // act(src) transform role
    const bool sclamp = S.clamp != 0, oclamp = a.out_clamp != 0;
    const double inv_n = 1.0 / (double)a.Rt;

    int cur_t = -1;
    for (int tile = tile_lo, it = 0; tile < tile_hi; ++tile, ++it) {
        const int buf = it % a.nbuf;
        const int t = tile / tps, r0 = (tile - t * tps) * kWgR, rows = min(kWgR, a.Rt - r0);
        if (t != cur_t) {
            __syncthreads();
            for (int j = tid; j < nw; j += 256) {
                int p, s, l, n;
                float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pw_col(d, pn * gwp + n0 + j, p, s, l, n)) {
                    const size_t idx = (size_t)t * a.cpo + s;
                    c = bnbwd_consts(a.tb[p].aff[idx], a.tb[p].bnp[idx], a.tb[p].bsum[idx], inv_n);
                }
                s_colc[tcol(j, NTW >> 3)] = c;
            }
            for (int i = tid; i < kw; i += 256) s_aff[tcol(i, kWgK >> 3)] = S.aff ? S.aff[(size_t)t * S.cp + k0 + i] : make_float2(1.f, 0.f);
            cur_t = t;
            __syncthreads();
        }
        mbar_wait(&full[buf], (it / a.nbuf) & 1);
        const unsigned char* rb = raw + (size_t)buf * L.raw_stride;
        if (have_dr) {
            if (rrl < rnrl) {                            // plain copy of this CTA's dR columns
                const uint4* dv = reinterpret_cast<const uint4*>(rb);
                const int nch = NP >> 3, ch = ((pn * gwp + n0) >> 3) + rq;
                for (int r = rrl; r < kWgR; r += rnrl)
                    *reinterpret_cast<uint4*>(Rs + (size_t)r * ldr + rq * 8) = r < rows ? dv[r * nch + ch] : make_uint4(0, 0, 0, 0);
            }
        } else if (rrl < rnrl) {
            float4 c8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q] = s_colc[q * (NTW >> 3) + rq];
            const uint4* dv = reinterpret_cast<const uint4*>(rb + (size_t)kWgR * o_dout);
            const uint4* ov = reinterpret_cast<const uint4*>(rb + (size_t)kWgR * o_out);
            if (a.direct) {
                dv = reinterpret_cast<const uint4*>(a.dout[pn] + ((size_t)t * a.Rt + r0) * a.cpo);
                ov = reinterpret_cast<const uint4*>(a.out[pn] + ((size_t)t * a.Rt + r0) * a.cpo);
            }
            const int nch = a.cpo >> 3, ch = (n0 >> 3) + rq;
            for (int r = rrl; r < kWgR; r += rnrl) {
                uint4 dvv = make_uint4(0, 0, 0, 0);
                if (r < rows) {
                    dvv = dv[r * nch + ch]; const uint4 ovv = ov[r * nch + ch];
                    uint32_t* dw = reinterpret_cast<uint32_t*>(&dvv); const uint32_t* ow = reinterpret_cast<const uint32_t*>(&ovv);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 dd = unpack2(dw[i]), oo = unpack2(ow[i]);
                        dw[i] = pack2(bnbwd_apply(dd.x, oo.x, c8[2 * i], oclamp), bnbwd_apply(dd.y, oo.y, c8[2 * i + 1], oclamp));
                    }
                }
                *reinterpret_cast<uint4*>(Rs + (size_t)r * ldr + rq * 8) = dvv;       // rows past the slice end contribute zero
            }
        }
        if (xrl < xnrl) {
            float2 c8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q] = s_aff[q * (kWgK >> 3) + xq];
            const uint4* sv = reinterpret_cast<const uint4*>(rb + (size_t)kWgR * o_src);
            const int nch = S.cp >> 3, ch = (k0 >> 3) + xq;
            for (int r = xrl; r < kWgR; r += xnrl) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (r < rows) v = affine8(sv[r * nch + ch], c8, sclamp);
                *reinterpret_cast<uint4*>(Xs + (size_t)r * ldx + xq * 8) = v;
            }
        }
        __syncthreads();
        {
            const int mi = lane >> 3;
#pragma unroll
            for (int ks = 0; ks < kWgR; ks += 16) {
                uint32_t af[4];
                ldsm4t(af, Xs + (size_t)(ks + (mi >> 1) * 8 + (lane & 7)) * ldx + kg * 16 + (mi & 1) * 8);
#pragma unroll
                for (int nb2 = 0; nb2 < NBW / 2; ++nb2) {
                    uint32_t bfr[4];
                    ldsm4t(bfr, Rs + (size_t)(ks + (mi & 1) * 8 + (lane & 7)) * ldr + nh * NBW * 8 + nb2 * 16 + (mi >> 1) * 8);
                    mma16816(acc[2 * nb2], af, bfr[0], bfr[1]);
                    mma16816(acc[2 * nb2 + 1], af, bfr[2], bfr[3]);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && tile + a.nbuf < tile_hi) issue(tile + a.nbuf, buf);
    }
    // ---- accumulate this CTA's partial tile into the gradient arena (logical layout)
    {
        const int g = lane >> 2, tg = lane & 3;
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kk = koff + k0 + kg * 16 + g + (e >> 1) * 8;
                const int j = pn * gwp + n0 + nh * NBW * 8 + nb * 8 + 2 * tg + (e & 1);
                int lk, k, p, s, lj, n;
                if (kg * 16 + g + (e >> 1) * 8 < kw && nh * NBW * 8 + nb * 8 + 2 * tg + (e & 1) < nw &&
                    pw_row(d, kk, lk, k) && pw_col(d, j, p, s, lj, n) && lk == lj)
                    atomicAdd(d.layer[lk].dw + (size_t)k * d.layer[lk].N + n, acc[nb][e]);
            }
        }
    }
    // ---- BatchNorm parameter gradients (one CTA): dgamma = sum_t S2, dbeta = sum_t S1
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        for (int j = tid; j < NP; j += 256) {
            int p, s, l, n;
            if (!pw_col(d, j, p, s, l, n)) continue;
            double gs = 0.0, bs = 0.0;
            for (int t = 0; t < kT; ++t) { const double2 v = a.tb[p].bsum[(size_t)t * a.cpo + s]; bs += v.x; gs += v.y; }
            d.layer[l].dg[n] = (float)gs; d.layer[l].dbe[n] = (float)bs;
        }
    }
}

// ======================================================================================== weight gradient on tcgen05
// The same product on the 5th-generation tensor cores: dW[kk][j] = sum_r act(src)[r][kk] * dR[r][j] with BOTH operands
// MN-major (the reduction runs over rows), staged by the transform passes straight into the 128-byte-swizzled layout,
// and the WHOLE [KP x NPall] accumulator resident in tensor memory for the life of the CTA: every CTA covers all K and
// N of its rows, so each row tile is loaded and transformed exactly once (the mma.sync kernel above re-reads it once per
// (K tile, N tile)).  One elected thread issues tcgen05.mma (M = 128 per K block, N = NPall, K = 16 rows); staging is
// double buffered against the tensor pipe through tcgen05.commit -> mbarrier.  Epilogue: TMEM -> registers -> shared
// (transposed) -> coalesced fp32 atomics in the reference's logical layout.
struct PwWgTcSmem { int colc, aff, rmap, cmap, raw, stg, total, raw_stride, stg_stride, xs_bytes, mb, np, nblk, tmem_cols; };
inline __host__ __device__ PwWgTcSmem pw_wgrad_tc_smem(int R, int KP, int NPall, int nplanes, int cpo, int src_cp_sum, int nbuf) {
    PwWgTcSmem s;
    s.mb = (KP + 127) / 128; s.np = (NPall + 15) & ~15; s.nblk = (s.np + 63) / 64;
    int cols = s.mb * s.np; s.tmem_cols = 32; while (s.tmem_cols < cols) s.tmem_cols *= 2;
    int off = 128;                                     // mbarriers (8 TMA + 2 MMA) + TMEM slot
    s.colc = off; off += s.np * 16;
    s.aff = off; off += s.mb * 128 * 8;
    s.rmap = off; off += s.mb * 128 * 4;
    s.cmap = off; off += s.np * 4;
    off = (off + 1023) & ~1023;
    s.xs_bytes = s.mb * 2 * R * 128;
    s.stg_stride = s.xs_bytes + s.nblk * R * 128;      // [Xs blocks | Rs blocks], multiple of 1024 for R >= 8
    s.stg = off; off += 2 * s.stg_stride;
    { const int scratch = 128 * 65 * 4; if (2 * s.stg_stride < scratch) off = s.stg + ((scratch + 1023) & ~1023); }   // epilogue transposition tile
    (void)nplanes; (void)cpo;
    s.raw_stride = (R * (NPall + src_cp_sum) * 2 + 127) & ~127;       // [dR rows | source rows]
    s.raw = off; off += nbuf * s.raw_stride;
    s.total = off + 1024;                              // slack for the manual 1024-byte alignment of the base
    return s;
}

template <int R, int NT>
__global__ void __launch_bounds__(NT, 1) pw_wgrad_tc_kernel(const PwBwdArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment for the swizzled tiles by OFFSET (an integer round trip of the pointer would drop the shared
    // address space and turn every tile access into a generic load)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const PwDesc& d = *a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    const int NP = d.NPall, gwp = d.cols.gwp, nplanes = d.cols.nplanes, KP = d.KP;
    int src_cp_sum = 0;
    for (int i = 0; i < d.nsrc; ++i) src_cp_sum += d.src[i].cp;
    const PwWgTcSmem L = pw_wgrad_tc_smem(R, KP, NP, nplanes, a.cpo, src_cp_sum, a.nbuf);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [0..7] TMA ring, [8..9] MMA done
    uint64_t* mma_done = full + 8;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + 96);
    float4* s_colc = reinterpret_cast<float4*>(smem + L.colc);
    float2* s_aff = reinterpret_cast<float2*>(smem + L.aff);
    int* s_rmap = reinterpret_cast<int*>(smem + L.rmap);
    int* s_cmap = reinterpret_cast<int*>(smem + L.cmap);
    unsigned char* raw = smem + L.raw;
    const int o_src = NP * 2;                          // per-row byte offset of the source regions inside a raw buffer

    const int tps = (a.Rt + R - 1) / R, ntile = kT * tps;
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(ntile, tile_lo + a.tiles_per_cta);
    if (warp == 0) tmem_alloc(s_tmem, (uint32_t)L.tmem_cols);
    if (tid == 0) { for (int b = 0; b < 10; ++b) mbar_init(&full[b], 1); mbar_fence_init(); }
    // logical maps: GEMM row kk -> (layer << 24 | k), GEMM column j -> (layer << 24 | n); -1 = padding
    for (int kk = tid; kk < L.mb * 128; kk += NT) { int l, k; s_rmap[kk] = (kk < KP && pw_row(d, kk, l, k)) ? ((l << 24) | k) : -1; }
    for (int j = tid; j < L.np; j += NT) { int p, sl, l, n; s_cmap[j] = (j < NP && pw_col(d, j, p, sl, l, n)) ? ((l << 24) | n) : -1; }
    for (int i = tid; i < 2 * L.stg_stride / 16; i += NT) reinterpret_cast<uint4*>(smem + L.stg)[i] = make_uint4(0, 0, 0, 0);   // padding rows / columns stay zero
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t idesc = umma_idesc(128, L.np, 1, 1);

    auto issue = [&](int tile, int buf) {
        const int t = tile / tps, r0 = (tile - t * tps) * R, rows = min(R, a.Rt - r0);
        unsigned char* dst = raw + (size_t)buf * L.raw_stride;
        const size_t row = (size_t)t * a.Rt + r0;
        mbar_expect_tx(&full[buf], (uint32_t)rows * (NP + src_cp_sum) * 2);
        bulk_g2s(dst, a.dr + row * NP, rows * NP * 2, &full[buf]);
        int off = 0;
        for (int i = 0; i < d.nsrc; ++i) {
            bulk_g2s(dst + (size_t)R * (o_src + off), d.src[i].data + row * d.src[i].cp, rows * d.src[i].cp * 2, &full[buf]);
            off += d.src[i].cp * 2;
        }
    };
    pdl_wait();
    if (tid == 32) for (int b = 0; b < a.nbuf; ++b) if (tile_lo + b < tile_hi) issue(tile_lo + b, b);     // deep ring: nbuf - 1 tiles in flight

    // transform roles: thread <-> one 8-column chunk (fixed), row lanes stride the rows
    const int nqr = NP >> 3, rq = tid % nqr, rrl = tid / nqr, rnrl = NT / nqr;                    // dR chunks (plain copy)
    const int nqx = src_cp_sum >> 3, xq = tid % nqx, xrl = tid / nqx, xnrl = NT / nqx;            // act(src) chunks over all sources
    int xsrc = 0, xch = xq, xoffb = 0;                                                              // source of this thread's chunk
    while (xsrc < d.nsrc - 1 && xch >= (d.src[xsrc].cp >> 3)) { xch -= d.src[xsrc].cp >> 3; xoffb += d.src[xsrc].cp * 2; ++xsrc; }
    const int xnch = d.src[xsrc].cp >> 3;
    const bool sclamp = d.src[xsrc].clamp != 0;

    int cur_t = -1;
    float2 c8[8];
    int nissued[2] = {0, 0};                           // commits per staging set (thread 0 bookkeeping is CTA-uniform)
    for (int tile = tile_lo, it = 0; tile < tile_hi; ++tile, ++it) {
        const int buf = it % a.nbuf, sb = it & 1;
        const int t = tile / tps, r0 = (tile - t * tps) * R, rows = min(R, a.Rt - r0);
        if (t != cur_t) {
            __syncthreads();
            {
                int off = 0;
                for (int i = 0; i < d.nsrc; ++i) {
                    for (int k = tid; k < d.src[i].cp; k += NT) s_aff[off + k] = d.src[i].aff ? d.src[i].aff[(size_t)t * d.src[i].cp + k] : make_float2(1.f, 0.f);
                    off += d.src[i].cp;
                }
            }
            cur_t = t;
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q] = s_aff[xq * 8 + q];     // this thread's chunk constants for the whole slice
        }
        mbar_wait(&full[buf], (it / a.nbuf) & 1);
        if (nissued[sb] > 0) { mbar_wait(&mma_done[sb], (nissued[sb] - 1) & 1); tc_fence_after(); }      // staging set free again
        unsigned char* Xs = smem + L.stg + (size_t)sb * L.stg_stride;
        unsigned char* Rs = Xs + L.xs_bytes;
        const unsigned char* rb = raw + (size_t)buf * L.raw_stride;
        if (rrl < rnrl) {                                  // dR rows -> swizzled MN-major tile (rows past the slice end contribute zero)
            const uint4* dv = reinterpret_cast<const uint4*>(rb);
#pragma unroll 4
            for (int r = rrl; r < R; r += rnrl)
                *reinterpret_cast<uint4*>(Rs + sw128_offset(r, rq * 8, R)) = r < rows ? dv[r * nqr + rq] : make_uint4(0, 0, 0, 0);
        }
        if (xrl < xnrl) {
            const uint4* sv = reinterpret_cast<const uint4*>(rb + (size_t)R * (o_src + xoffb));
#pragma unroll 2
            for (int r = xrl; r < R; r += xnrl) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (r < rows) v = affine8(sv[r * xnch + xch], c8, sclamp);
                *reinterpret_cast<uint4*>(Xs + sw128_offset(r, xq * 8, R)) = v;
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 32 && tile + a.nbuf < tile_hi) issue(tile + a.nbuf, buf);
        if (tid == 0) {
            tc_fence_after();
            const uint32_t xa = smem_u32(Xs), ra = smem_u32(Rs);
            for (int mb = 0; mb < L.mb; ++mb)
#pragma unroll
                for (int ks = 0; ks < R / 16; ++ks)
                    umma_bf16(tmem + (uint32_t)(mb * L.np), umma_desc(xa + mb * 2 * R * 128 + ks * 2048, R * 128, 1024),
                              umma_desc(ra + ks * 2048, R * 128, 1024), idesc, it > 0 || ks > 0);
            umma_commit(&mma_done[sb]);
        }
        ++nissued[sb];
    }
    // ---- drain: the last commit of each staging set covers every MMA issued before it
    for (int sb = 0; sb < 2; ++sb)
        if (nissued[sb] > 0) mbar_wait(&mma_done[sb], (nissued[sb] - 1) & 1);
    tc_fence_after();
    __syncthreads();
    // ---- epilogue: 64-column chunks through a transposition tile, then coalesced atomics (logical layout)
    if (tile_lo < tile_hi) {
        float* S = reinterpret_cast<float*>(smem + L.stg);          // [128][65]
        for (int mb = 0; mb < L.mb; ++mb)
            for (int c0 = 0; c0 < L.np; c0 += 64) {
                const int ncol = min(64, L.np - c0);
                if (warp < 4) {
                    for (int c = 0; c < ncol; c += 8) {
                        float v[8];
                        tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(mb * L.np + c0 + c), v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) S[(32 * warp + lane) * 65 + c + i] = v[i];
                    }
                }
                __syncthreads();
                for (int i = tid; i < 128 * ncol; i += NT) {
                    const int row = i / ncol, col = i - row * ncol;
                    const int rm = s_rmap[mb * 128 + row], cm = s_cmap[c0 + col];
                    if (rm >= 0 && cm >= 0 && (rm >> 24) == (cm >> 24)) {
                        const LayerP& Lp = d.layer[rm >> 24];
                        atomicAdd(Lp.dw + (size_t)(rm & 0xffffff) * Lp.N + (cm & 0xffffff), S[row * 65 + col]);
                    }
                }
                __syncthreads();
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)L.tmem_cols);
    // ---- BatchNorm parameter gradients (one CTA): dgamma = sum_t S2, dbeta = sum_t S1
    if (blockIdx.x == 0) {
        for (int j = tid; j < NP; j += NT) {
            int p, sl, l, n;
            if (!pw_col(d, j, p, sl, l, n)) continue;
            double gs = 0.0, bs = 0.0;
            for (int t = 0; t < kT; ++t) { const double2 v = a.tb[p].bsum[(size_t)t * a.cpo + sl]; bs += v.x; gs += v.y; }
            d.layer[l].dg[n] = (float)gs; d.layer[l].dbe[n] = (float)bs;
        }
    }
}

// ======================================================================================== depthwise backward
// One frame at a time per CTA: d out + raw out frames by TMA (dense), input frame rows by TMA into the halo-padded
// tile (activated in place); dR (BatchNorm backward of the depthwise output) goes into a second halo-padded tile.
// thread <-> (channel pair, column) runs the 9-tap weight gradient (registers, reduced once per CTA) and the
// transposed stencil (data gradient + the BatchNorm-backward sums of the input tensor).  The sums are taken from the
// ACTIVATED input a = relu6(scale*raw + shift) held in shared memory: mask = 0 < a < 6, xhat = (a - beta) / gamma.
template <int CP, int S>
__global__ void __launch_bounds__(kDwThreads) dw_bwd_kernel(const DwArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NPAIR = CP / 2, NCH = CP / 8, NXL = kDwThreads / NPAIR, TNPL = kDwThreads / NCH;
    const int tid = threadIdx.x;
    pdl_trigger();
    const int in_px = a.Hi * a.Wi, out_px = a.Ho * a.Wo, PW = a.Wi + 2, QW = a.Wo + 2;
    const DwSmem L = dw_smem(CP, a.Hi, a.Wi, a.Ho, a.Wo, a.nbuf, true, a.band_rows, S);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    float* s_stat = reinterpret_cast<float*>(smem + L.stat);
    float* s_wred = reinterpret_cast<float*>(smem + L.wred);
    float4* s_colc = reinterpret_cast<float4*>(smem + L.colc);
    bf16* Pin = reinterpret_cast<bf16*>(smem + L.pin);
    bf16* Pdr = reinterpret_cast<bf16*>(smem + L.pdr);
    const uint32_t orow_bytes = (uint32_t)a.Wo * CP * 2, row_bytes = (uint32_t)a.Wi * CP * 2;
    const int nitems = kT * a.B * a.nbands;
    const int it_lo = blockIdx.x * a.frames_per_cta, it_hi = min(nitems, it_lo + a.frames_per_cta);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    for (int i = tid; i < CP * 2; i += kDwThreads) s_stat[i] = 0.f;
    for (int i = tid; i < CP * 9; i += kDwThreads) s_wred[i] = 0.f;
    for (int i = tid; i < L.pin_stride / 4; i += kDwThreads) reinterpret_cast<uint32_t*>(Pin)[i] = 0u;
    for (int i = tid; i < (a.band_rows + 2) * QW * (CP / 2); i += kDwThreads) reinterpret_cast<uint32_t*>(Pdr)[i] = 0u;
    __syncthreads();
    // rows of (d out, out) a band needs: its own and one neighbour row on either side (the data gradient of the band's
    // input rows reaches them)
    auto out_rows = [&](const DwBand& b, int& o_lo, int& o_hi) { o_lo = max(0, b.oy0 - 1); o_hi = min(a.Ho - 1, b.oy0 + b.bho); };
    auto issue = [&](int item, int buf) {
        const DwBand b = dw_band(a, item, S);
        int o_lo, o_hi; out_rows(b, o_lo, o_hi);
        const uint32_t ob = orow_bytes * (uint32_t)(o_hi - o_lo + 1), ib = row_bytes * (uint32_t)(b.i_hi - b.i_lo + 1);
        mbar_expect_tx(&full[buf], 2 * ob + ib);
        bulk_g2s(smem + L.raw_dout + (size_t)buf * L.out_stride, a.dout + ((size_t)b.f * out_px + (size_t)o_lo * a.Wo) * CP, ob, &full[buf]);
        bulk_g2s(smem + L.raw_out + (size_t)buf * L.out_stride, a.out + ((size_t)b.f * out_px + (size_t)o_lo * a.Wo) * CP, ob, &full[buf]);
        bulk_g2s(smem + L.rawin + (size_t)buf * L.in_stride, a.in + ((size_t)b.f * in_px + (size_t)b.i_lo * a.Wi) * CP, ib, &full[buf]);
    };
    pdl_wait();
    if (tid == 0) for (int b = 0; b < a.nbuf; ++b) if (it_lo + b < it_hi) issue(it_lo + b, b);

    const int tch = tid % NCH, tpl = tid / NCH;
    const int pr = tid % NPAIR, xl = tid / NPAIR;
    const bool active = xl < NXL;
    float w0[9], w1[9], g0[9], g1[9];
    {
        const int l0 = slot_logical(a.map, 2 * pr), l1 = slot_logical(a.map, 2 * pr + 1);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            w0[k] = l0 >= 0 ? a.L.w[k * a.L.N + a.kbase + l0] : 0.f;
            w1[k] = l1 >= 0 ? a.L.w[k * a.L.N + a.kbase + l1] : 0.f;
            g0[k] = 0.f; g1[k] = 0.f;
        }
    }
    float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
    float4 xc0 = make_float4(1.f, 0.f, 0.f, 0.f), xc1 = xc0;    // (scale, shift, inv_std, -mean * inv_std) of this thread's two input channels
    float2 ac8[8];
    const bool iclamp = a.clamp != 0, xform = a.aff != nullptr || iclamp, want_sums = a.in_bsum != nullptr, banded = a.nbands > 1;
    const double inv_n = 1.0 / ((double)a.B * out_px);
    auto flush = [&](int t) {
        if (active) {
            atomicAdd(&s_stat[4 * pr], s1a); atomicAdd(&s_stat[4 * pr + 1], s2a);
            atomicAdd(&s_stat[4 * pr + 2], s1b); atomicAdd(&s_stat[4 * pr + 3], s2b);
        }
        s1a = s2a = s1b = s2b = 0.f;
        __syncthreads();
        for (int c = tid; c < CP; c += kDwThreads) {
            if (want_sums && c >= a.in_sum_lo && c < a.in_sum_hi) {
                double2* dst = a.in_bsum + (size_t)t * CP + c;
                atomicAdd(&dst->x, (double)s_stat[2 * c]); atomicAdd(&dst->y, (double)s_stat[2 * c + 1]);
            }
            s_stat[2 * c] = 0.f; s_stat[2 * c + 1] = 0.f;
        }
        __syncthreads();
    };
    auto xhat_consts = [&](int t, int slot) { return sum_consts(a.aff, a.bnp, (size_t)t * CP + slot); };
    const int row_step = S * PW * CP;
    int cur_t = -1;
    for (int item = it_lo, it = 0; item < it_hi; ++item, ++it) {
        const DwBand bd = dw_band(a, item, S);
        int o_lo, o_hi; out_rows(bd, o_lo, o_hi);
        const int buf = it % a.nbuf, f = bd.f, t = f / a.B;
        if (t != cur_t) {
            if (cur_t >= 0) flush(cur_t);
            __syncthreads();
            for (int s = tid; s < CP; s += kDwThreads) {
                const int l = slot_logical(a.map, s);
                const size_t idx = (size_t)t * CP + s;
                // chunk-transposed (constant q of every 8-slot chunk contiguous): the 16-byte reads below are conflict free
                s_colc[(s & 7) * NCH + (s >> 3)] = l >= 0 ? bnbwd_consts(a.tb.aff[idx], a.tb.bnp[idx], a.tb.bsum[idx], inv_n) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (tpl < TNPL) {
#pragma unroll
                for (int q = 0; q < 8; ++q) ac8[q] = a.aff ? a.aff[(size_t)t * CP + tch * 8 + q] : make_float2(1.f, 0.f);
            }
            xc0 = xhat_consts(t, 2 * pr); xc1 = xhat_consts(t, 2 * pr + 1);
            cur_t = t;
            __syncthreads();
        }
        mbar_wait(&full[buf], (it / a.nbuf) & 1);
        const bf16* Rin = reinterpret_cast<const bf16*>(smem + L.rawin + (size_t)buf * L.in_stride);
        if (tpl < TNPL) {
            float4 c8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) c8[q] = s_colc[q * NCH + tch];
            const uint4* dv = reinterpret_cast<const uint4*>(smem + L.raw_dout + (size_t)buf * L.out_stride) + tch;
            const uint4* ov = reinterpret_cast<const uint4*>(smem + L.raw_out + (size_t)buf * L.out_stride) + tch;
            uint4* qv = reinterpret_cast<uint4*>(Pdr) + tch;
            {   // dR rows o_lo .. o_hi -> tile rows (oy - oy0 + 1)
                PxWalk wo(tpl, TNPL, a.Wo);
                const int npx = (o_hi - o_lo + 1) * a.Wo, roff = o_lo - bd.oy0 + 1;
                for (int px = tpl; px < npx; px += TNPL, wo.next()) {
                    uint4 dvv = dv[px * NCH]; const uint4 ovv = ov[px * NCH];
                    uint32_t* dw = reinterpret_cast<uint32_t*>(&dvv); const uint32_t* ow = reinterpret_cast<const uint32_t*>(&ovv);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 dd = unpack2(dw[i]), oo = unpack2(ow[i]);
                        dw[i] = pack2(bnbwd_apply(dd.x, oo.x, c8[2 * i], false), bnbwd_apply(dd.y, oo.y, c8[2 * i + 1], false));
                    }
                    qv[((wo.y + roff) * QW + wo.x + 1) * NCH] = dvv;
                }
                if (banded) {                               // halo rows outside the frame must read as zero
                    if (bd.oy0 == 0) for (int x = tpl; x < QW; x += TNPL) qv[x * NCH] = make_uint4(0, 0, 0, 0);
                    if (bd.oy0 + bd.bho >= a.Ho) for (int x = tpl; x < QW; x += TNPL) qv[((bd.bho + 1) * QW + x) * NCH] = make_uint4(0, 0, 0, 0);
                }
            }
            {   // raw input rows i_lo .. i_hi -> activated halo tile, tile row (iy - row0)
                uint4* pv = reinterpret_cast<uint4*>(Pin) + tch;
                const uint4* rv = reinterpret_cast<const uint4*>(Rin) + tch;
                PxWalk wi(tpl, TNPL, a.Wi);
                const int npx = (bd.i_hi - bd.i_lo + 1) * a.Wi, roff = bd.i_lo - bd.row0;
                for (int px = tpl; px < npx; px += TNPL, wi.next()) {
                    const uint4 v = rv[px * NCH];
                    pv[((wi.y + roff) * PW + wi.x + 1) * NCH] = xform ? affine8(v, ac8, iclamp) : v;
                }
                if (banded) {
                    const int th = (bd.bho - 1) * S + 3;
                    for (int tr = 0; tr < th; ++tr) {
                        const int iy = bd.row0 + tr;
                        if (iy >= 0 && iy < a.Hi) continue;
                        for (int x = tpl; x < PW; x += TNPL) pv[(tr * PW + x) * NCH] = make_uint4(0, 0, 0, 0);
                    }
                }
            }
        }
        __syncthreads();
        if (active) {
            // weight gradient over the band's own output rows: dw[ky][kx] += act(in)(oy*S - pt + ky, ox*S - pl + kx) * dR(oy, ox)
            // (rows shared between consecutive outputs stay in registers)
            for (int ox = xl; ox < a.Wo; ox += NXL) {
                const bf16* win = Pin + (ox * S + 1 - a.pad_l) * CP + 2 * pr;
                const bf16* dp = Pdr + (QW + ox + 1) * CP + 2 * pr;
                auto emit = [&](const float2 (&A)[3], const float2 (&B)[3], const float2 (&C)[3]) {
                    const float2 dr = unpack2(*reinterpret_cast<const uint32_t*>(dp));
                    dw_grad_row(A, dr, g0, g1); dw_grad_row(B, dr, g0 + 3, g1 + 3); dw_grad_row(C, dr, g0 + 6, g1 + 6);
                    dp += QW * CP;
                };
                float2 A[3], B[3], C[3];
                if (S == 1) {
                    dw_ldrow<CP>(win, A); dw_ldrow<CP>(win + PW * CP, B);
                    const bf16* nxt = win + 2 * PW * CP;
                    for (int oy = 0; oy < bd.bho; oy += 3) {
                        dw_ldrow<CP>(nxt, C); emit(A, B, C); nxt += PW * CP;
                        if (oy + 1 < bd.bho) { dw_ldrow<CP>(nxt, A); emit(B, C, A); nxt += PW * CP; }
                        if (oy + 2 < bd.bho) { dw_ldrow<CP>(nxt, B); emit(C, A, B); nxt += PW * CP; }
                    }
                } else {
                    dw_ldrow<CP>(win, A);
                    for (int oy = 0; oy < bd.bho; ++oy) {
                        dw_ldrow<CP>(win + PW * CP, B); dw_ldrow<CP>(win + 2 * PW * CP, C);
                        emit(A, B, C);
#pragma unroll
                        for (int j = 0; j < 3; ++j) A[j] = C[j];
                        win += row_step;
                    }
                }
            }
            // data gradient of the input rows the band owns: [oy0 * S, (oy0 + bho) * S) (clipped to the frame)
            //     d in(iy, ix) = sum_{ky,kx} w[ky][kx] * dR((iy + pt - ky)/S, (ix + pl - kx)/S)
            const int own_lo = bd.oy0 * S, own_hi = min(a.Hi, (bd.oy0 + bd.bho) * S);
            for (int ix = xl; ix < a.Wi; ix += NXL) {
                bf16* gp = a.din + ((size_t)f * in_px + (size_t)own_lo * a.Wi + ix) * CP + 2 * pr;
                // the sums use the RAW input value (same mask / xhat as every consumer of them, bnbwd_apply): the ring holds it
                const bf16* rp = Rin + ((size_t)(own_lo - bd.i_lo) * a.Wi + ix) * CP + 2 * pr;
                // (+ existing share: `ex`, loaded by the caller one row ahead), store, BatchNorm-backward sums of the input
                auto finish = [&](float acc0, float acc1, uint32_t ex) {
                    if (a.accumulate) { const float2 e = unpack2(ex); acc0 += e.x; acc1 += e.y; }
                    const uint32_t pk = pack2(acc0, acc1);
                    *reinterpret_cast<uint32_t*>(gp) = pk;
                    if (want_sums) {
                        const float2 gr = unpack2(pk);
                        const float2 rv = unpack2(*reinterpret_cast<const uint32_t*>(rp));
                        sum_accum(gr.x, rv.x, xc0, iclamp, s1a, s2a);
                        sum_accum(gr.y, rv.y, xc1, iclamp, s1b, s2b);
                    }
                    gp += a.Wi * CP; rp += a.Wi * CP;
                };
                if (S == 1) {
                    // own row li = iy - oy0 sees dR rows iy - 1 .. iy + 1 = tile rows li .. li + 2 <-> ky = 2, 1, 0
                    auto emit = [&](const float2 (&A)[3], const float2 (&B)[3], const float2 (&C)[3]) {
                        float acc0 = 0.f, acc1 = 0.f;
                        dw_mac_row<true>(A, w0 + 6, w1 + 6, acc0, acc1);
                        dw_mac_row<true>(B, w0 + 3, w1 + 3, acc0, acc1);
                        dw_mac_row<true>(C, w0, w1, acc0, acc1);
                        finish(acc0, acc1, a.accumulate ? *reinterpret_cast<const uint32_t*>(gp) : 0u);
                    };
                    const int nown = own_hi - own_lo;
                    const bf16* dwin = Pdr + ix * CP + 2 * pr;
                    float2 A[3], B[3], C[3];
                    dw_ldrow<CP>(dwin, A); dw_ldrow<CP>(dwin + QW * CP, B);
                    const bf16* nxt = dwin + 2 * QW * CP;
                    for (int li = 0; li < nown; li += 3) {
                        dw_ldrow<CP>(nxt, C); emit(A, B, C); nxt += QW * CP;
                        if (li + 1 < nown) { dw_ldrow<CP>(nxt, A); emit(B, C, A); nxt += QW * CP; }
                        if (li + 2 < nown) { dw_ldrow<CP>(nxt, B); emit(C, A, B); nxt += QW * CP; }
                    }
                } else {
                    // stride 2: input pixel (iy, ix) only sees the taps whose parity matches, over even (iy+pt-ky), (ix+pl-kx).
                    // The column taps are fixed per thread (kx = 0, 2 at tile columns cA, cA - 1 when ix + pl is even; kx = 1 at
                    // cA otherwise) and selected ONCE; the row taps alternate with the parity of iy + pt.  dR row n = tile row n - oy0 + 1.
                    const int xs = ix + a.pad_l;
                    const bool xe = (xs & 1) == 0;
                    const int cA = (xs >> 1) + 1, cB = xe ? cA - 1 : cA;
                    float wa0[3], wa1[3], wb0[3], wb1[3];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        wa0[ky] = xe ? w0[ky * 3] : w0[ky * 3 + 1]; wa1[ky] = xe ? w1[ky * 3] : w1[ky * 3 + 1];
                        wb0[ky] = xe ? w0[ky * 3 + 2] : 0.f;         wb1[ky] = xe ? w1[ky * 3 + 2] : 0.f;
                    }
                    const bf16* pa = Pdr + cA * CP + 2 * pr; const bf16* pb = Pdr + cB * CP + 2 * pr;
                    // the existing share of row iy + 1 is requested before row iy is computed (a dependent global load per row otherwise)
                    uint32_t ex_cur = (a.accumulate && own_lo < own_hi) ? *reinterpret_cast<const uint32_t*>(gp) : 0u;
                    for (int iy = own_lo; iy < own_hi; ++iy) {
                        const uint32_t ex_next = (a.accumulate && iy + 1 < own_hi) ? *reinterpret_cast<const uint32_t*>(gp + a.Wi * CP) : 0u;
                        const int ys = iy + a.pad_t;
                        float acc0, acc1;
                        if ((ys & 1) == 0) {                      // ky = 0 at dR row ys/2, ky = 2 at dR row ys/2 - 1
                            const int r = ((ys >> 1) - bd.oy0 + 1) * QW * CP;
                            const float2 a0 = unpack2(*reinterpret_cast<const uint32_t*>(pa + r)), b0 = unpack2(*reinterpret_cast<const uint32_t*>(pb + r));
                            const float2 a2 = unpack2(*reinterpret_cast<const uint32_t*>(pa + r - QW * CP)), b2 = unpack2(*reinterpret_cast<const uint32_t*>(pb + r - QW * CP));
                            acc0 = fmaf(a0.x, wa0[0], fmaf(b0.x, wb0[0], fmaf(a2.x, wa0[2], b2.x * wb0[2])));
                            acc1 = fmaf(a0.y, wa1[0], fmaf(b0.y, wb1[0], fmaf(a2.y, wa1[2], b2.y * wb1[2])));
                        } else {                                  // ky = 1 at dR row (ys - 1)/2
                            const int r = (((ys - 1) >> 1) - bd.oy0 + 1) * QW * CP;
                            const float2 a1 = unpack2(*reinterpret_cast<const uint32_t*>(pa + r)), b1 = unpack2(*reinterpret_cast<const uint32_t*>(pb + r));
                            acc0 = fmaf(a1.x, wa0[1], b1.x * wb0[1]);
                            acc1 = fmaf(a1.y, wa1[1], b1.y * wb1[1]);
                        }
                        finish(acc0, acc1, ex_cur);
                        ex_cur = ex_next;
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && item + a.nbuf < it_hi) issue(item + a.nbuf, buf);
    }
    if (cur_t >= 0) flush(cur_t);
    // ---- weight gradients of this CTA
    if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { atomicAdd(&s_wred[(2 * pr) * 9 + k], g0[k]); atomicAdd(&s_wred[(2 * pr + 1) * 9 + k], g1[k]); }
    }
    __syncthreads();
    for (int i = tid; i < CP * 9; i += kDwThreads) {
        const int s = i / 9, k = i - s * 9, l = slot_logical(a.map, s);
        if (l >= 0) atomicAdd(a.L.dw + (size_t)k * a.L.N + a.kbase + l, s_wred[i]);
    }
    if (blockIdx.x == 0) {
        for (int s = tid; s < CP; s += kDwThreads) {
            const int l = slot_logical(a.map, s);
            if (l < 0) continue;
            double gs = 0.0, bs = 0.0;
            for (int t = 0; t < kT; ++t) { const double2 v = a.tb.bsum[(size_t)t * CP + s]; bs += v.x; gs += v.y; }
            a.L.dg[a.kbase + l] = (float)gs; a.L.dbe[a.kbase + l] = (float)bs;
        }
    }
}

// ======================================================================================== global average pool backward
// d head[f][p][c] = d gap[f][c] / HW (gradient wrt the activated head output) + its BatchNorm-backward sums
__global__ void __launch_bounds__(256) gap_bwd_kernel(const GapArgs a, int frames_per_block) {
    extern __shared__ float s_acc[];              // [cp][2]
    const int t = blockIdx.y, nch = a.cp >> 3, tid = threadIdx.x;
    for (int i = tid; i < a.cp * 2; i += 256) s_acc[i] = 0.f;
    __syncthreads();
    const int b_lo = blockIdx.x * frames_per_block, b_hi = min(a.B, b_lo + frames_per_block);
    const float inv = 1.0f / (float)a.HW;
    for (int item = tid; item < (b_hi - b_lo) * nch; item += 256) {
        const int b = b_lo + item / nch, ch = item % nch, f = t * a.B + b;
        float4 c8[8];
        float gv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            c8[q] = sum_consts(a.aff, a.bnp, (size_t)t * a.cp + ch * 8 + q);
            gv[q] = ch * 8 + q < a.C ? a.dgap[(size_t)f * a.C + ch * 8 + q] * inv : 0.f;
        }
        uint4 gvec;
        uint32_t* gw = reinterpret_cast<uint32_t*>(&gvec);
#pragma unroll
        for (int i = 0; i < 4; ++i) gw[i] = pack2(gv[2 * i], gv[2 * i + 1]);
        float s1[8], s2[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { s1[q] = 0.f; s2[q] = 0.f; }
        const uint4* src = reinterpret_cast<const uint4*>(a.in + (size_t)f * a.HW * a.cp) + ch;
        uint4* dst = reinterpret_cast<uint4*>(a.dout + (size_t)f * a.HW * a.cp) + ch;
        for (int p = 0; p < a.HW; ++p) {
            const uint4 rv = src[(size_t)p * nch];
            dst[(size_t)p * nch] = gvec;
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 x = unpack2(gw[i]), y = unpack2(rw[i]);
                sum_accum(x.x, y.x, c8[2 * i], true, s1[2 * i], s2[2 * i]);
                sum_accum(x.y, y.y, c8[2 * i + 1], true, s1[2 * i + 1], s2[2 * i + 1]);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) { atomicAdd(&s_acc[(ch * 8 + q) * 2], s1[q]); atomicAdd(&s_acc[(ch * 8 + q) * 2 + 1], s2[q]); }
    }
    __syncthreads();
    for (int c = tid; c < a.cp; c += 256) {
        double2* d = a.bsum + (size_t)t * a.cp + c;
        atomicAdd(&d->x, (double)s_acc[2 * c]); atomicAdd(&d->y, (double)s_acc[2 * c + 1]);
    }
}

}  // namespace v2
}  // namespace cdra
#endif
