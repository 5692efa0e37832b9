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
    int frames_per_cta;
};

__global__ void __launch_bounds__(kDwThreads) dw_fwd_kernel(const DwArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int cp = a.cp, npair = cp >> 1;
    const int in_px = a.Hi * a.Wi, out_px = a.Ho * a.Wo;
    const uint32_t frame_bytes = (uint32_t)in_px * cp * 2;
    const int buf_stride = (frame_bytes + 127) & ~127;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    float* s_stat = reinterpret_cast<float*>(smem + 64);                 // [cp][2]
    unsigned char* fb = smem + ((64 + cp * 8 + 127) & ~127);
    const int nframes = kT * a.B;
    const int f_lo = blockIdx.x * a.frames_per_cta, f_hi = min(nframes, f_lo + a.frames_per_cta);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    for (int i = tid; i < cp * 2; i += kDwThreads) s_stat[i] = 0.f;
    __syncthreads();
    auto issue = [&](int f, int buf) {
        mbar_expect_tx(&full[buf], frame_bytes);
        bulk_g2s(fb + (size_t)buf * buf_stride, a.in + (size_t)f * in_px * cp, frame_bytes, &full[buf]);
    };
    if (tid == 0) {
        if (f_lo < f_hi) issue(f_lo, 0);
        if (f_lo + 1 < f_hi) issue(f_lo + 1, 1);
    }
    // transform role: fixed 8-slot chunk, pixel lanes
    const int nch = cp >> 3, tch = tid % nch, tpl = tid / nch, tnpl = kDwThreads / nch;
    float2 c8[8];
    // stencil role: fixed channel pair, column lanes
    const int pr = tid % npair, xl = tid / npair, nxl = kDwThreads / npair;
    const bool active = xl < nxl;
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
    float ssum0 = 0.f, ssum1 = 0.f, ssq0 = 0.f, ssq1 = 0.f;
    auto flush = [&](int t) {                         // CTA-uniform
        if (active) {
            atomicAdd(&s_stat[4 * pr], ssum0); atomicAdd(&s_stat[4 * pr + 1], ssq0);
            atomicAdd(&s_stat[4 * pr + 2], ssum1); atomicAdd(&s_stat[4 * pr + 3], ssq1);
        }
        ssum0 = ssum1 = ssq0 = ssq1 = 0.f;
        __syncthreads();
        for (int c = tid; c < cp; c += kDwThreads) {
            double2* dst = a.tb.fsum + (size_t)t * cp + c;
            atomicAdd(&dst->x, (double)s_stat[2 * c]); atomicAdd(&dst->y, (double)s_stat[2 * c + 1]);
            s_stat[2 * c] = 0.f; s_stat[2 * c + 1] = 0.f;
        }
        __syncthreads();
    };

    int cur_t = -1;
    for (int f = f_lo, it = 0; f < f_hi; ++f, ++it) {
        const int buf = it & 1, t = f / a.B;
        if (t != cur_t) {
            if (cur_t >= 0 && a.training) flush(cur_t);
            if (tpl < tnpl) {
#pragma unroll
                for (int q = 0; q < 8; ++q) c8[q] = a.aff ? a.aff[(size_t)t * cp + tch * 8 + q] : make_float2(1.f, 0.f);
            }
            cur_t = t;
        }
        mbar_wait(&full[buf], (it >> 1) & 1);
        bf16* fr = reinterpret_cast<bf16*>(fb + (size_t)buf * buf_stride);
        if (a.aff != nullptr || a.clamp) {            // producer's BatchNorm affine (+ReLU6), in place
            if (tpl < tnpl) {
                uint4* v = reinterpret_cast<uint4*>(fr);
                for (int px = tpl; px < in_px; px += tnpl) v[px * nch + tch] = affine8(v[px * nch + tch], c8, a.clamp != 0);
            }
            __syncthreads();
        }
        if (active) {
            for (int ox = xl; ox < a.Wo; ox += nxl) {
                const int ix0 = ox * a.stride - a.pad_l;
                bf16* ocol = a.out + ((size_t)f * out_px + ox) * cp + 2 * pr;
                for (int oy = 0; oy < a.Ho; ++oy) {
                    const int iy0 = oy * a.stride - a.pad_t;
                    float acc0 = b0, acc1 = b1;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int iy = iy0 + ky;
                        if (iy < 0 || iy >= a.Hi) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int ix = ix0 + kx;
                            if (ix < 0 || ix >= a.Wi) continue;
                            const float2 v = unpack2(*reinterpret_cast<const uint32_t*>(fr + ((size_t)iy * a.Wi + ix) * cp + 2 * pr));
                            acc0 = fmaf(v.x, w0[ky * 3 + kx], acc0);
                            acc1 = fmaf(v.y, w1[ky * 3 + kx], acc1);
                        }
                    }
                    const uint32_t pk = pack2(acc0, acc1);
                    *reinterpret_cast<uint32_t*>(ocol + (size_t)oy * a.Wo * cp) = pk;
                    const float2 r = unpack2(pk);
                    ssum0 += r.x; ssq0 = fmaf(r.x, r.x, ssq0);
                    ssum1 += r.y; ssq1 = fmaf(r.y, r.y, ssq1);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && f + 2 < f_hi) issue(f + 2, buf);
    }
    if (cur_t >= 0 && a.training) flush(cur_t);
    if (a.counter == nullptr) return;
    if (!last_cta(a.counter, gridDim.x)) return;
    for (int s = tid; s < cp; s += kDwThreads) {
        const int l = slot_logical(a.map, s);
        if (l >= 0) bn_finalize_channel(a.tb, cp, s, a.L, a.kbase + l, (double)a.B * out_px, a.training);
        else for (int t = 0; t < kT; ++t) { a.tb.aff[(size_t)t * cp + s] = make_float2(0.f, 0.f); a.tb.bnp[(size_t)t * cp + s] = make_float2(0.f, 1.f); }
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
