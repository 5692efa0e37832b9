// v2 tower: depthwise 3x3 convolutions (stride 1 | 2, TF SAME padding; core/architectures.py:132,138), the global
// average pool (:172) and their backward passes.  A frame of a tower tensor ([H*W][cp] bf16) is contiguous in HBM:
// a CTA pulls whole frames into shared memory with one TMA bulk copy each (double buffered), applies the producer's
// BatchNorm affine (+ReLU6) in place with 16-byte vectors, and runs the stencil with thread <-> (channel pair,
// output column): weights live in registers, shared-memory reads are conflict-free (consecutive lanes = consecutive
// channel pairs), output stores are coalesced 4-byte pairs, BatchNorm sums accumulate in registers.
#pragma once
#ifndef CDRA_EMU
#include "v2_common.cuh"

namespace cdra {
namespace v2 {

constexpr int kDwThreads = 512;

struct DwArgs {
    const bf16* in; const float2* aff; const float2* bnp; int clamp;     // input tensor [4*B*Hi*Wi][cp]
    bf16* din; double2* in_bsum; int in_sum_lo, in_sum_hi; int accumulate; // backward: gradient wrt the activated input
    int cp; SlotMap map; int kbase;                                      // slot -> layer channel = kbase + logical(slot)
    int B, Hi, Wi, Ho, Wo, stride, pad_t, pad_l;
    LayerP L;                                                            // w [9][C], b, g, be [C]
    bf16* out; const bf16* dout; Tables tb;                              // output tensor [4*B*Ho*Wo][cp] (raw) + tables
    int training;
    unsigned* counter;
    int frames_per_cta, nbuf;                                            // work items (frame, band) per CTA; depth of the TMA ring
    int band_rows, nbands;                                               // output rows per band; bands per frame (1: whole frames)
};

// shared-memory carve-up of the depthwise kernels (host + device).  The unit of work is a BAND of `bh` output rows of one
// frame (the whole frame when it fits): the activated input rows the band needs live in a halo-padded tile
// [(bh - 1) * S + 3][(Wi + 2)][cp] whose row 0 is input row oy0 * S - pad_t (rows / columns outside the frame are zero = TF
// SAME padding), so the stencils need no bounds checks.  Forward: the TMA row copies land in the tile's interior and the
// producer's BatchNorm affine (+ReLU6) is applied in place.  Backward: the ring holds the RAW input rows (one contiguous
// bulk copy per band; the BatchNorm-backward sums of the input need the raw values) next to the rows oy0 - 1 .. oy0 + bh of
// (d out, out); the activated input tile and the dR tile (with a one-row halo of REAL neighbour rows) exist once.
struct DwSmem { int stat, wred, colc, pin, raw_out, raw_dout, pdr, total, pin_stride, out_stride, rawin, in_stride, th; };
inline __host__ __device__ DwSmem dw_smem(int cp, int Hi, int Wi, int Ho, int Wo, int nbuf, bool backward, int bh, int S) {
    DwSmem s;
    int off = 64;
    s.stat = off; off += cp * 8;
    s.wred = off; off += backward ? cp * 9 * 4 : 0;
    s.colc = off; off += backward ? cp * 16 : 0;
    off = (off + 127) & ~127;
    s.th = (bh - 1) * S + 3;
    const int in_rows = s.th < Hi ? s.th : Hi;          // input rows a band loads at most
    s.pin_stride = (s.th * (Wi + 2) * cp * 2 + 127) & ~127;
    s.in_stride = (in_rows * Wi * cp * 2 + 127) & ~127;
    s.out_stride = ((bh + 2 < Ho ? bh + 2 : Ho) * Wo * cp * 2 + 127) & ~127;
    s.pin = off; off += (backward ? 1 : nbuf) * s.pin_stride;
    s.rawin = off; off += backward ? nbuf * s.in_stride : 0;
    s.raw_out = off; off += backward ? nbuf * s.out_stride : 0;
    s.raw_dout = off; off += backward ? nbuf * s.out_stride : 0;
    s.pdr = off; off += backward ? (((bh + 2) * (Wo + 2) * cp * 2 + 127) & ~127) : 0;   // dR rows oy0 - 1 .. oy0 + bh
    s.total = off;
    return s;
}
// one work item: frame f, output rows [oy0, oy0 + bho); tile row 0 = input row `row0`; the band loads input rows [i_lo, i_hi]
struct DwBand { int f, oy0, bho, row0, i_lo, i_hi; };
CDRA_DEV DwBand dw_band(const DwArgs& a, int item, int S) {
    DwBand b;
    b.f = item / a.nbands;
    b.oy0 = (item - b.f * a.nbands) * a.band_rows;
    b.bho = min(a.band_rows, a.Ho - b.oy0);
    b.row0 = b.oy0 * S - a.pad_t;
    b.i_lo = max(0, b.row0);
    b.i_hi = min(a.Hi - 1, b.row0 + (b.bho - 1) * S + 2);
    return b;
}

// per-thread (y, x) walker over the pixels px = lane, lane + step, ... of a W-wide frame without divisions in the loop
struct PxWalk {
    int y, x, dy, dx, W;
    __device__ PxWalk(int lane, int step, int W_) : W(W_) { y = lane / W_; x = lane - y * W_; dy = step / W_; dx = step - dy * W_; }
    __device__ void next() { y += dy; x += dx; if (x >= W) { x -= W; ++y; } }
};

// three horizontally adjacent taps (one channel pair each) of a tile row; the stencils slide DOWN a column and keep the rows
// they share between consecutive outputs in registers (stride 1: one new row per output instead of three)
template <int CP>
CDRA_DEV void dw_ldrow(const bf16* p, float2 (&r)[3]) {
    r[0] = unpack2(*reinterpret_cast<const uint32_t*>(p));
    r[1] = unpack2(*reinterpret_cast<const uint32_t*>(p + CP));
    r[2] = unpack2(*reinterpret_cast<const uint32_t*>(p + 2 * CP));
}
// acc += sum_j r[j] * w[j]   (REV: taps mirrored, w[2 - j])
template <bool REV>
CDRA_DEV void dw_mac_row(const float2 (&r)[3], const float* w0, const float* w1, float& a0, float& a1) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        a0 = fmaf(r[j].x, w0[REV ? 2 - j : j], a0);
        a1 = fmaf(r[j].y, w1[REV ? 2 - j : j], a1);
    }
}
CDRA_DEV void dw_grad_row(const float2 (&r)[3], float2 dr, float* g0, float* g1) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { g0[j] = fmaf(r[j].x, dr.x, g0[j]); g1[j] = fmaf(r[j].y, dr.y, g1[j]); }
}

// forward: out(oy, ox) = b + sum_{ky,kx} w[ky][kx] * act(in)(oy*S - pt + ky, ox*S - pl + kx)   (zero outside the frame)
template <int CP, int S>
__global__ void __launch_bounds__(kDwThreads, 2) dw_fwd_kernel(const DwArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NPAIR = CP / 2, NCH = CP / 8, NXL = kDwThreads / NPAIR, TNPL = kDwThreads / NCH;
    const int tid = threadIdx.x;
    pdl_trigger();
    const int in_px = a.Hi * a.Wi, out_px = a.Ho * a.Wo, PW = a.Wi + 2;
    const DwSmem L = dw_smem(CP, a.Hi, a.Wi, a.Ho, a.Wo, a.nbuf, false, a.band_rows, S);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    float* s_stat = reinterpret_cast<float*>(smem + L.stat);
    const uint32_t row_bytes = (uint32_t)a.Wi * CP * 2;
    const int nitems = kT * a.B * a.nbands;
    const int it_lo = blockIdx.x * a.frames_per_cta, it_hi = min(nitems, it_lo + a.frames_per_cta);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    for (int i = tid; i < CP * 2; i += kDwThreads) s_stat[i] = 0.f;
    for (int i = tid; i < a.nbuf * L.pin_stride / 4; i += kDwThreads) reinterpret_cast<uint32_t*>(smem + L.pin)[i] = 0u;
    __syncthreads();
    auto issue = [&](int item, int buf) {
        const DwBand b = dw_band(a, item, S);
        mbar_expect_tx(&full[buf], row_bytes * (uint32_t)(b.i_hi - b.i_lo + 1));
        unsigned char* dst = smem + L.pin + (size_t)buf * L.pin_stride + CP * 2;        // column 1 of tile row 0
        const bf16* src = a.in + (size_t)b.f * in_px * CP;
        for (int y = b.i_lo; y <= b.i_hi; ++y) bulk_g2s(dst + (size_t)(y - b.row0) * PW * CP * 2, src + (size_t)y * a.Wi * CP, row_bytes, &full[buf]);
    };
    pdl_wait();
    if (tid == 0) for (int b = 0; b < a.nbuf; ++b) if (it_lo + b < it_hi) issue(it_lo + b, b);

    const int tch = tid % NCH, tpl = tid / NCH;                       // transform role
    const int pr = tid % NPAIR, xl = tid / NPAIR;                     // stencil role
    const bool active = xl < NXL;
    float w0[9], w1[9], b0 = 0.f, b1 = 0.f;
    {
        const int l0 = slot_logical(a.map, 2 * pr), l1 = slot_logical(a.map, 2 * pr + 1);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            w0[k] = l0 >= 0 ? a.L.w[k * a.L.N + a.kbase + l0] : 0.f;
            w1[k] = l1 >= 0 ? a.L.w[k * a.L.N + a.kbase + l1] : 0.f;
        }
        if (l0 >= 0) b0 = a.L.b[a.kbase + l0];
        if (l1 >= 0) b1 = a.L.b[a.kbase + l1];
    }
    float2 c8[8];
    float ssum0 = 0.f, ssum1 = 0.f, ssq0 = 0.f, ssq1 = 0.f;
    auto flush = [&](int t) {
        if (active) {
            atomicAdd(&s_stat[4 * pr], ssum0); atomicAdd(&s_stat[4 * pr + 1], ssq0);
            atomicAdd(&s_stat[4 * pr + 2], ssum1); atomicAdd(&s_stat[4 * pr + 3], ssq1);
        }
        ssum0 = ssum1 = ssq0 = ssq1 = 0.f;
        __syncthreads();
        for (int c = tid; c < CP; c += kDwThreads) {
            double2* dst = a.tb.fsum + (size_t)t * CP + c;
            atomicAdd(&dst->x, (double)s_stat[2 * c]); atomicAdd(&dst->y, (double)s_stat[2 * c + 1]);
            s_stat[2 * c] = 0.f; s_stat[2 * c + 1] = 0.f;
        }
        __syncthreads();
    };
    const bool clamp = a.clamp != 0, xform = a.aff != nullptr || clamp, banded = a.nbands > 1;
    const int row_step = S * PW * CP, ostep = a.Wo * CP;
    int cur_t = -1;
    for (int item = it_lo, it = 0; item < it_hi; ++item, ++it) {
        const DwBand bd = dw_band(a, item, S);
        const int buf = it % a.nbuf, f = bd.f, t = f / a.B;
        if (t != cur_t) {
            if (cur_t >= 0 && a.training) flush(cur_t);
            if (tpl < TNPL) {
#pragma unroll
                for (int q = 0; q < 8; ++q) c8[q] = a.aff ? a.aff[(size_t)t * CP + tch * 8 + q] : make_float2(1.f, 0.f);
            }
            cur_t = t;
        }
        mbar_wait(&full[buf], (it / a.nbuf) & 1);
        bf16* Pin = reinterpret_cast<bf16*>(smem + L.pin + (size_t)buf * L.pin_stride);
        if (xform || banded) {
            if (tpl < TNPL) {
                uint4* pv = reinterpret_cast<uint4*>(Pin) + tch;
                if (xform) {                                          // producer's BatchNorm affine (+ReLU6), in place, on the loaded rows
                    PxWalk w(tpl, TNPL, a.Wi);
                    const int npx = (bd.i_hi - bd.i_lo + 1) * a.Wi, roff = bd.i_lo - bd.row0;
                    for (int px = tpl; px < npx; px += TNPL, w.next()) {
                        uint4* q = pv + ((w.y + roff) * PW + w.x + 1) * NCH;
                        *q = affine8(*q, c8, clamp);
                    }
                }
                if (banded) {                                         // tile rows outside the frame: a previous band left data there
                    const int th = (bd.bho - 1) * S + 3;
                    for (int tr = 0; tr < th; ++tr) {
                        const int iy = bd.row0 + tr;
                        if (iy >= 0 && iy < a.Hi) continue;
                        for (int x = tpl; x < PW; x += TNPL) pv[(tr * PW + x) * NCH] = make_uint4(0, 0, 0, 0);
                    }
                }
            }
            __syncthreads();
        }
        if (active) {
            for (int ox = xl; ox < a.Wo; ox += NXL) {
                const bf16* win = Pin + (ox * S + 1 - a.pad_l) * CP + 2 * pr;         // tile row 0 = tap ky 0 of the band's first row
                bf16* optr = a.out + ((size_t)f * out_px + (size_t)bd.oy0 * a.Wo + ox) * CP + 2 * pr;
                auto emit = [&](const float2 (&A)[3], const float2 (&B)[3], const float2 (&C)[3]) {
                    float acc0 = b0, acc1 = b1;
                    dw_mac_row<false>(A, w0, w1, acc0, acc1);
                    dw_mac_row<false>(B, w0 + 3, w1 + 3, acc0, acc1);
                    dw_mac_row<false>(C, w0 + 6, w1 + 6, acc0, acc1);
                    const uint32_t pk = pack2(acc0, acc1);
                    *reinterpret_cast<uint32_t*>(optr) = pk;
                    const float2 r = unpack2(pk);
                    ssum0 += r.x; ssq0 = fmaf(r.x, r.x, ssq0);
                    ssum1 += r.y; ssq1 = fmaf(r.y, r.y, ssq1);
                    optr += ostep;
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
        }
        __syncthreads();
        if (tid == 0 && item + a.nbuf < it_hi) issue(item + a.nbuf, buf);
    }
    if (cur_t >= 0 && a.training) flush(cur_t);
    if (a.counter == nullptr) return;
    if (!last_cta(a.counter, gridDim.x)) return;
    for (int s = tid; s < CP; s += kDwThreads) {
        const int l = slot_logical(a.map, s);
        if (l >= 0) bn_finalize_channel(a.tb, CP, s, a.L, a.kbase + l, (double)a.B * out_px, a.training);
        else for (int t = 0; t < kT; ++t) { a.tb.aff[(size_t)t * CP + s] = make_float2(0.f, 0.f); a.tb.bnp[(size_t)t * CP + s] = make_float2(0.f, 1.f); }
    }
}

// ---------------------------------------------------------------------------------------------- global average pool
struct GapArgs {
    const bf16* in; const float2* aff; int cp, C, HW, F;      // head conv raw [F*HW][cp]; F = 4*B frames
    int B;
    float* out;                                                // [F][C] fp32
    const float* dgap; bf16* dout;                             // backward: d out [F][C] -> d head (activated) [F*HW][cp]
    const float2* bnp; double2* bsum;
};

__global__ void __launch_bounds__(256) gap_fwd_kernel(const GapArgs a) {
    const int nch = a.cp >> 3;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (idx >= (long long)a.F * nch) return;
    const int f = (int)(idx / nch), ch = (int)(idx - (long long)f * nch), t = f / a.B;
    float2 c8[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) c8[q] = a.aff[(size_t)t * a.cp + ch * 8 + q];
    float s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = 0.f;
    const uint4* src = reinterpret_cast<const uint4*>(a.in + (size_t)f * a.HW * a.cp) + ch;
    for (int p = 0; p < a.HW; ++p) {
        const uint4 v = src[(size_t)p * nch];
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 x = unpack2(w[i]);
            s[2 * i] += relu6f(fmaf(x.x, c8[2 * i].x, c8[2 * i].y));
            s[2 * i + 1] += relu6f(fmaf(x.y, c8[2 * i + 1].x, c8[2 * i + 1].y));
        }
    }
    const float inv = 1.0f / (float)a.HW;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (ch * 8 + q < a.C) a.out[(size_t)f * a.C + ch * 8 + q] = s[q] * inv;
}

}  // namespace v2
}  // namespace cdra
#endif
